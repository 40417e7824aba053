"""vslam-pose-estimation-framework_b200: B200-native (sm_100a) stereo framepoint generation and frame-aligner
linearisation behind the plugin interfaces of Ssellu/vslam-pose-estimation-framework (ProSLAM fork).

The directory name is not a Python identifier; import it through the alias package `vslam_b200`
(vslam_b200/__init__.py), e.g. `from vslam_b200 import api, configs, synth`.

  csrc/      hand-written CUDA kernels + the C ABI (include/vslam_b200.h) -> libvslam_b200.so
  api.py     ctypes binding + Python mirror of the reference's generator / aligner interfaces
  configs.py effective parameter sets of the reference's configurations/*.yaml
  synth.py   seeded synthetic workloads (BASELINE.json configs)
"""
