"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU/PyTorch-fp32 restatement of the reference's per-step training path (MECLabTUDA/Lifelong-nnUNet @ fb55c48 and the
un-vendored nnunet@77bc485 pieces it calls).  The product (lifelong-nnunet_b200/) never imports this package; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and there only as the checker
or as the CPU baseline being timed.

Pinning status: the CL losses / POD / KD / EWC arithmetic is pinned against the reference's own unmodified files
(tests/test_oracle_vs_reference.py + tests/golden/*.npz); the network / Dice+CE arithmetic lives in un-vendored
nnunet and is pinned only structurally (reference test_MultiHead_Module.py fixture) -> "parity unpinned" there.
"""
