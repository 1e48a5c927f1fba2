"""TEST INFRASTRUCTURE ONLY.

CPU oracles for the JuicySuite hot path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (juicy-audio-plugins_b200/) never does.

  oracle.refhost   -- the reference's own C++ compiled unmodified (oracle/_ref/*.so)
  oracle.port      -- this repo's C restatement (oracle/liboracle_port.so)
"""
