"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's hot path (NumPy environments, Philox streams, SB3 PPO
semantics, the scalar reference port) and the script that made tests/golden/.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this package; nothing under three-mlagents_b200/ does."""
