import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        if l.strip():
            print(l.strip()[:200])
        continue
    print("%-40s %6d %-4s %9.2f ms %7.1f G ch-samples/s %6.1f%% of HBM" % ("+".join(x.replace("Juicy", "") for x in d["chain"]), d["clips"], d["path"],
          d["ms_per_render"], d["ch_samples_per_s"] / 1e9, 100 * d["frac_of_measured_hbm"]))
