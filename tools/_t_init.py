import sys
sys.path.insert(0, '.')
from vslam_b200 import api, configs, synth
cfg = configs.KITTI; cam = synth.camera(cfg.camera)
l, r = synth.band_world_pair(cfg.camera, 1)
gen = api.StereoFramePointGenerator(cfg, cam)
print(gen.initialize(l, r, True))
