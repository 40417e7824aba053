"""CPU oracle of the reference hot path -- TEST INFRASTRUCTURE ONLY (see oracle/c/vslam_oracle.h).

tier_a : C restatement of every stage (no OpenCV), the bit-exact spec of the CUDA kernels.
tier_b : the same pipeline with OpenCV 4.13 (python cv2) for the three primitives the reference calls
         from OpenCV; cross-check of tier_a and the CPU baseline bench.py times.
ref    : ctypes binding of oracle/_ref/libvslam_ref.so -- the reference's own, unmodified hot-path translation units
         compiled against the functional third-party stand-ins of oracle/shims (oracle/Makefile `_ref`): what tier_a is
         pinned to (tests/test_oracle_vs_ref.py), and the CPU arm bench.py times (kind "reference").
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
