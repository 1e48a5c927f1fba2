for m in 0 1 2 3 4 8 12 7 15; do
  echo -n "skip=$m "; JB_COOP_DEBUG_SKIP=$m python bench.py --no-cpu --steps 3 2>/dev/null | grep -o "mean_launch_ms[^,]*"
done
