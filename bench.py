#!/usr/bin/env python3
"""bench.py -- stereo frames/s of the hot path (framepoint generation + aligner linearize) on N B200s.

Workload (BASELINE.json configs[2], "configuration_kitti_fast.yaml, 4096 independent KITTI-shape stereo pairs
batched and sharded across 1/2/4/8 B200"): every rank owns one GPU and a batch of `--pairs` independent 1241x376
pairs, each processed as a first frame (status Localizing).  One step = one pass over the rank's batch:

    initialize (FAST + NMS, border filter, blur, rBRIEF)  ->  compute (epipolar Hamming scan, bins, triangulation)
    ->  StereoUVAligner initialize + `--rounds` x linearize per pair over the pair's new framepoints

value   : whole-job frames/s with the images already resident in HBM (CUDA events, max over ranks)
e2e     : the same through the C-ABI call that takes HOST buffers (H2D of the images and D2H of the framepoint
          records and normal equations inside the timed region)
roofline: dominant kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
cpu_baseline / --impl reference: the oracle pipeline with OpenCV primitives (tier B) on the host cores.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torch.distributed.run (one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stereo frames/s (framepoint gen + aligner linearize)"
CONFIG_NAME = "kitti_fast"
KERNEL_OF_INTEREST = "fast_nms"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=4096, help="stereo pairs per rank per step")
    ap.add_argument("--distinct", type=int, default=0,
                    help="distinct synthetic pairs generated per rank (0: 256 for weak scaling, every pair for strong)")
    ap.add_argument("--rounds", type=int, default=10, help="linearize rounds per pair per step (BASELINE config 4)")
    ap.add_argument("--cpu-sample", type=int, default=1024,
                    help="pairs of the single-thread CPU baseline sample (at most; the sample also ends after ~14 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --pairs per GPU; strong: --pairs in total, block-partitioned over the ranks "
                         "(BASELINE.json configs[2]: 4096 pairs sharded across 1/2/4/8)")
    ap.add_argument("--wc-images", action="store_true", help="image upload buffers in write-combined pinned memory")
    ap.add_argument("--no-extras", action="store_true", help="skip the aligner / sequence / landmark side measurements")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# CPU side (oracle; the checker timed as the reference's CPU path -- never on the product path)
# ------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(left, right, rounds):
    _CPU.update(left=left, right=right, rounds=rounds)


def _cpu_kind():
    """"reference": oracle/_ref -- the reference's own unmodified translation units (stereo_framepoint_generator.cpp,
    stereouv_aligner.cpp, ...) with OpenCV's own FAST / ORB (python cv2) behind its cv:: calls; "port": the tier-B
    restatement, when the prebuilt library did not travel"""
    if "kind" not in _CPU:
        from oracle import ref
        _CPU["kind"] = "reference" if ref.available() else "port"
    return _CPU["kind"]


def _cpu_pair(i, as_configured=False):
    """one stereo pair through the reference's CPU path (OpenCV primitives, 1 thread): returns n framepoints.
    as_configured: also run the FLANN knnMatch + findHomography block that `use_matches: true` executes and whose
    results the reference never reads (stereo_framepoint_generator.cpp:168-273) -- tier B only"""
    from oracle import pipeline, tier_a
    from vslam_b200 import configs, synth
    cfg, acfg = configs.BY_NAME[CONFIG_NAME], configs.ALIGNER_BY_NAME[CONFIG_NAME]
    cam = synth.camera(cfg.camera)
    if _cpu_kind() == "reference" and not as_configured:
        from oracle import ref
        if "session" not in _CPU or _CPU.get("pid") != os.getpid():
            _CPU["cv2"] = ref.install_cv2_backend()
            _CPU["session"] = ref.Session(cam, None, **ref.effective_values(CONFIG_NAME)).configure()
            _CPU["pid"] = os.getpid()
            _CPU["T"] = _prior()
        return _CPU["session"].first_frame(_CPU["left"][i], _CPU["right"][i], _CPU["rounds"], _CPU["T"])
    if "gen" not in _CPU:
        _CPU["gen"] = pipeline.StereoFramePointGeneratorOracle(cfg, cam, "b")
        _CPU["T"] = _prior()
    gen = _CPU["gen"]
    gen.thresholds[:] = cfg.detector_threshold_minimum       # every pair is a first frame
    gen.initialize(_CPU["left"][i], _CPU["right"][i], True)
    if as_configured:
        gen.dead_use_matches_block()
    gen.compute()
    fp = gen.framepoints()
    moving = np.ascontiguousarray(fp["cam"])
    fixed = np.stack([fp["xl"], fp["yl"], fp["xr"], fp["yr"]], 1).astype(np.float64)
    wt = np.minimum(acfg.maximum_reliable_depth_meters / moving[:, 2], 1.0) if len(fp) else np.zeros(0)
    t0 = time.perf_counter()
    al = tier_a.Aligner("stereouv", moving, fixed, np.ones(len(fp)), wt, cam.K, cam.baseline, cam.rows, cam.cols,
                        acfg.minimum_reliable_depth_meters, acfg.maximum_error_kernel)
    for _ in range(_CPU["rounds"]):
        al.linearize(_CPU["T"], False)
    _CPU["pose_optimization"] = _CPU.get("pose_optimization", 0.0) + time.perf_counter() - t0   # pose_tracker_3d.h:123-127
    return len(fp)


def _prior():
    from vslam_b200 import synth
    return synth.true_motion(0.15)


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _cpu_what(rounds):
    import cv2
    if _cpu_kind() == "reference":
        return ("oracle/_ref: the reference's own unmodified StereoFramePointGenerator::initialize + compute and "
                "StereoUVAligner::linearize x%d (compiled from /root/reference against the stand-ins of oracle/shims), "
                "cv::FastFeatureDetector / cv::ORB answered by OpenCV %s (python cv2, cv2.setNumThreads(0) as "
                "executables/app.cpp:96); the dead use_matches FLANN block (stereo_framepoint_generator.cpp:168-273) "
                "returns empty results" % (rounds, cv2.__version__))
    return ("oracle tier B (cv2 %s FAST/ORB with cv2.setNumThreads(0) as executables/app.cpp:96, C -O2 stereo scan / bins / "
            "triangulation / linearize x%d); the reference's dead use_matches FLANN block is not executed"
            % (cv2.__version__, rounds))


def cpu_baseline(left, right, rounds, sample, budget_s=14.0):
    import cv2
    cv2.setNumThreads(0)
    _cpu_init(left, right, rounds)
    _cpu_pair(0)                                               # warm-up (imports, first-touch)
    kind = _cpu_kind()
    if kind == "reference":
        s = _CPU["session"]
        g0, p0 = s.generator_seconds(), s.seconds_pose_optimization
    else:
        _CPU["gen"].seconds = {k: 0.0 for k in _CPU["gen"].seconds}
        _CPU["pose_optimization"] = 0.0
    t0 = time.perf_counter()
    n = 0
    while n < sample and time.perf_counter() - t0 < budget_s:     # a bounded sample: `sample` pairs or `budget_s` seconds
        _cpu_pair(n % len(left))
        n += 1
    dt = time.perf_counter() - t0
    if kind == "reference":                                     # the reference's own chronometers (definitions.h:147-151)
        stages = {k: (v - g0[k]) / n * 1e3 for k, v in s.generator_seconds().items()}
        stages["pose_optimization"] = (s.seconds_pose_optimization - p0) / n * 1e3
    else:
        stages = {k: v / n * 1e3 for k, v in _CPU["gen"].seconds.items()}
        stages["pose_optimization"] = _CPU["pose_optimization"] / n * 1e3
    n2 = min(12, len(left))    # the as-configured variant (tier B + FLANN block) is several times slower: a smaller sample
    t0 = time.perf_counter()
    for i in range(n2):
        _cpu_pair(i, as_configured=True)
    dt2 = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": kind, "stages_ms_per_frame": stages,
            "host": {"cpu_model": _cpu_model(), "nproc": len(os.sched_getaffinity(0))},
            "as_configured": {"value": n2 / dt2, "unit": "frames/s", "sample": "%d pairs" % n2, "kind": "port",
                              "what": "oracle tier B plus the FLANN knnMatch(k=2) + findHomography(RANSAC) block that "
                                      "use_matches: true (struct default) executes and never reads "
                                      "(stereo_framepoint_generator.cpp:168-273)"},
            "sample": "%d of the same KITTI-shape first-frame pairs, one thread: %s" % (n, _cpu_what(rounds)),
            "seconds": dt}


def cpu_sequence_baseline(frames, name="kitti"):
    """BASELINE.json configs[0]: configuration_kitti.yaml, one synthetic 1241x376 stereo sequence through the tracker
    that executables/app runs per frame, on the CPU reference path: the reference's own PoseTracker3D::compute
    (initialize -> track -> StereoUVAligner -> prune -> recoverPoints -> landmark updates -> compute) from oracle/_ref,
    OpenCV's FAST / ORB behind it, one thread.  None when oracle/_ref did not travel."""
    from oracle import ref
    from vslam_b200 import configs, synth
    if not ref.available():
        return None
    version = ref.install_cv2_backend()
    cam = synth.camera(configs.BY_NAME[name].camera)
    s = ref.Session(cam, None, **ref.effective_values(name)).configure()
    t0 = time.perf_counter()
    for left, right in frames:
        s.process(left, right)
    dt = time.perf_counter() - t0
    stages = {k: v / len(frames) * 1e3 for k, v in {**s.generator_seconds(), **s.tracker_seconds()}.items()}
    pose = s.pose()
    out = {"value": len(frames) / dt, "unit": "frames/s", "cores": 1, "kind": "reference", "frames": len(frames),
           "stages_ms_per_frame": stages, "final_x_m": float(pose[0, 3]),
           "true_final_x_m": (len(frames) - 1) * (-cam.bx / cam.fx) / 4,
           "what": "oracle/_ref: the reference's unmodified PoseTracker3D::compute per frame (configuration_%s.yaml values), "
                   "cv::FastFeatureDetector / cv::ORB answered by OpenCV %s, one thread" % (name, version)}
    s.close()
    return out


def gpu_sequence_through_the_reference_tracker(frames, name="kitti"):
    """the same sequence through the same unmodified PoseTracker3D with adapters/ GpuStereoFramePointGenerator +
    GpuStereoUVAligner in place of the CPU classes (oracle/_ref/libvslam_ref_gpu.so, tests/test_gpu_dropin.py): what a
    user of the reference gets per frame after the two-line change of INTEGRATION.md, object graph and landmark
    bookkeeping of the reference included"""
    from oracle import ref
    from vslam_b200 import configs, synth
    if not ref.gpu_available():
        return None
    cam = synth.camera(configs.BY_NAME[name].camera)
    warm = ref.Session(cam, None, gpu=True, **ref.effective_values(name)).configure()
    for left, right in frames[:4]:      # module load, first allocations
        warm.process(left, right)
    warm.close()
    s = ref.Session(cam, None, gpu=True, **ref.effective_values(name)).configure()
    t0 = time.perf_counter()
    for left, right in frames:
        s.process(left, right)
    dt = time.perf_counter() - t0
    pose = s.pose()
    stages = {k: v / len(frames) * 1e3 for k, v in s.tracker_seconds().items()}
    out = {"value": len(frames) / dt, "unit": "frames/s", "frames": len(frames), "final_x_m": float(pose[0, 3]),
           "tracker_stages_ms_per_frame": stages,
           "what": "PoseTracker3D::compute of the reference (unmodified, oracle/_ref) driving adapters/ on the GPU"}
    s.close()
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref; the tier-B port only where the
    prebuilt library is missing) on every host core, one process per core, on OUR arm's config: each step is a bounded
    sample of that workload (4 pairs per core instead of --pairs), frames/s being independent of the batch size on a CPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    from vslam_b200 import configs, synth
    cfg = configs.BY_NAME[CONFIG_NAME]
    cores = len(os.sched_getaffinity(0))
    per_step = max(cores * 4, 32)
    distinct = min(args.distinct or 64, per_step)
    left, right = synth.band_world_batch(cfg.camera, range(distinct), workers=min(cores, 32))
    idx = [i % distinct for i in range(per_step)]
    _cpu_init(left, right, args.rounds)
    kind = _cpu_kind()
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_pair, idx, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_pair, idx, chunksize=1)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
            "config": workload_config(args, args.pairs),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                             "host": {"cpu_model": _cpu_model(), "nproc": cores},
                             "sample": "%d pairs per step (%d distinct) of the %d-pair workload, one process per core (%d): %s"
                                       % (per_step, distinct, args.pairs, cores, _cpu_what(args.rounds))},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_RESULT, flush=True)
    return 0


def workload_config(args, pairs):
    c = {"workload": "configuration_kitti_fast.yaml: independent KITTI-shape 1241x376 stereo pairs as first frames "
                     "(Localizing), FAST thr 15, bin 25, ORB-256, triangulation distance 25.6; + StereoUV "
                     "linearize x%d per pair over the pair's framepoints" % args.rounds,
         "linearize_rounds": args.rounds, "image": "1241x376 u8"}
    if args.scaling == "strong":
        c["pairs_total_per_step"] = pairs
        c["partition"] = "contiguous blocks of the %d pairs over the ranks (sharding.block_partition)" % pairs
        c["l2"] = "inputs larger than L2 (%.0f MB of images per step in total)" % (pairs * 2 * 1241 * 376 / 1e6)
    else:
        c["pairs_per_gpu_per_step"] = pairs
        c["l2"] = "inputs larger than L2 (%.0f MB of images per step per GPU)" % (pairs * 2 * 1241 * 376 / 1e6)
    return c


def bind_to_gpu_numa_node(torch, local_rank):
    """Best effort: run this rank (and hence first-touch its pinned host buffers) on the CPU cores NVML reports as local to
    its GPU, so that the H2D stream of `e2e` does not cross the socket interconnect when 4 or 8 ranks share the host.
    Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return "bound to %d of %d cores local to the GPU" % (len(cpus), len(allowed))
        return "all %d allowed cores are local to the GPU" % len(allowed)
    except Exception as e:   # no NVML, no permission: the run is still valid, only possibly slower end to end
        return "not bound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons of one GPU (B200_PROFILING.md recipe).  Started before the warm-up so that
    nvidia-smi's start-up latency does not eat the timed region; only samples that arrive between begin() and end()
    count (if the region is shorter than three sampling periods, every sample under load since start is used)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.08)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        inside = [x for x in self.lines if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30) + 0.06]
        window = "timed region"
        if len(inside) < 3:
            inside, window = self.lines, "warm-up + timed region (region shorter than 3 sampling periods)"
        sm, mx, reasons = parse(inside)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def aligner_stress(api, configs, synth, torch, dev):
    """BASELINE.json configs[3]: 100k synthetic 3D-2D correspondences, 10 Gauss-Newton rounds on one GPU, plus the
    linearize kernel alone on 4M correspondences (81 algorithmic bytes per correspondence per round)."""
    acfg = configs.KITTI_FAST_ALIGNER
    cam = synth.camera("kitti")
    out = {}
    peak, _ = peaks()
    for kind, cls in (("stereouv", api.StereoUVAligner), ("uvd", api.UVDAligner)):
        n = 100000
        c = synth.correspondences(n, kind, cam)
        omega = c["omega"] if kind == "stereouv" else np.stack([c["omega_uv"], c["omega_d"]], 1)
        al = cls(acfg, max_points=4_000_000)
        T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
        al.initialize(c["moving"], c["fixed"], omega, c["wt"], cam.K, cam.baseline, cam.rows, cam.cols, T0)
        stream = torch.cuda.ExternalStream(al.stream, device=dev)
        for _ in range(3):
            al.setPreviousToCurrent(T0)
            for _ in range(10):
                al.oneRound(False)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            al.setPreviousToCurrent(T0)
            for _ in range(10):
                al.oneRound(False)                      # host-driven: kernel + read-back + 6x6 solve per round
        stepwise_ms = (time.perf_counter() - t0) / reps * 1e3
        al.setPreviousToCurrent(T0)
        al.converge(fused=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            al.setPreviousToCurrent(T0)
            al.converge(fused=True)                     # one persistent kernel for the whole converge()
        fused_ms = (time.perf_counter() - t0) / reps * 1e3
        rounds = al.number_of_rounds
        # the linearize kernel alone at a size that exceeds L2: CUDA events on the library's stream
        big = 4_000_000
        reps_n = big // n
        big_set = (np.tile(c["moving"], (reps_n, 1)), np.tile(c["fixed"], (reps_n, 1)),
                   np.tile(omega, (reps_n, 1)) if omega.ndim == 2 else np.tile(omega, reps_n), np.tile(c["wt"], reps_n))
        al.initialize(*big_set, cam.K, cam.baseline, cam.rows, cam.cols, synth.true_motion())
        for _ in range(3):
            al.linearize_async(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        al.synchronize()
        e0.record(stream)
        for _ in range(10):
            al.linearize_async(False)
        e1.record(stream)
        al.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = 81.0 * big / (ms * 1e-3) / 1e9
        # size sweep of the linearize kernel alone (show the asymptote), uploads outside the timed launches
        sweep = []
        for m in (10_000, 100_000, 1_000_000, 4_000_000):
            al.initialize(*(a[:m] for a in big_set), cam.K, cam.baseline, cam.rows, cam.cols, synth.true_motion())
            for _ in range(3):
                al.linearize_async(False)
            al.synchronize()
            e0.record(stream)
            for _ in range(10):
                al.linearize_async(False)
            e1.record(stream)
            al.synchronize()
            ms_m = e0.elapsed_time(e1) / 10
            sweep.append({"n": m, "us": ms_m * 1e3, "gbs": 81.0 * m / (ms_m * 1e-3) / 1e9})
        out[kind] = {"n": n, "ten_rounds_host_driven_ms": stepwise_ms, "converge_fused_ms": fused_ms,
                     "converge_rounds": rounds, "fused_ms_per_round": fused_ms / max(rounds, 1),
                     "linearize_4M_ms": ms, "linearize_4M_gbs": gbs, "linearize_4M_frac_of_hbm": gbs / peak,
                     "linearize_sweep": sweep}
        al.close()
    return out


def landmark_refinement(api, synth, device):
    """Landmark::update for all tracked landmarks of a frame as one call (host buffers in and out,
    copies inside the timed region); 40 algorithmic bytes per measurement per Gauss-Newton iteration"""
    out = []
    for n, frames in ((1000, 60), (20000, 100)):
        h = synth.landmark_histories(n, n_frames=frames, seed=5, outlier_fraction=0.05)
        opt = api.LandmarkOptimizer(n, int(h["offsets"][-1]), frames, device=device)
        a = (h["offsets"], h["measurements"], h["world_to_camera"], h["camera_to_world"], h["world"], h["number_of_updates"])
        r = opt.update(*a)
        t0 = time.perf_counter()
        for _ in range(10):
            r = opt.update(*a)
        ms = (time.perf_counter() - t0) / 10 * 1e3
        out.append({"landmarks": n, "measurements": int(h["offsets"][-1]), "ms_per_call": ms,
                    "mean_iterations": float(r[3].mean()), "adopted_fraction": float((r[2] == 1).mean())})
        opt.close()
    # the device-resident map (vslam_landmark_map): histories stay in HBM, a frame appends one measurement per tracked
    # landmark and refines it -- what a tracker pays per frame for 1000 tracked landmarks with ~30-measurement histories
    n, frames = 1000, 60
    h = synth.landmark_histories(n, n_frames=frames, seed=5, outlier_fraction=0.05)
    lengths = np.diff(h["offsets"])
    lmap = api.LandmarkMap(n, 8 * n, frames + 64, device=device)
    for f in range(frames):     # poses of the history frames
        lmap.set_frame_pose(f, h["world_to_camera"][f], h["camera_to_world"][f])
    # every landmark is born with its history minus the last measurement (newest first), then updated once per "frame"
    born = [h["measurements"][h["offsets"][i]:h["offsets"][i + 1] - 1][::-1] for i in range(n)]
    born = [bm if len(bm) else h["measurements"][h["offsets"][i]:h["offsets"][i + 1]] for i, bm in enumerate(born)]
    offs = np.concatenate([[0], np.cumsum([len(bm) for bm in born])]).astype(np.int32)
    last = frames - 1
    r0 = lmap.update_frame(last, h["world_to_camera"][last], h["camera_to_world"][last], [], [], offs, np.concatenate(born), h["world"])
    ids = r0["new_ids"]
    cam = np.array([h["measurements"][h["offsets"][i + 1] - 1]["camera_coordinates"] for i in range(n)])
    lmap.update_frame(last, h["world_to_camera"][last], h["camera_to_world"][last], ids, cam)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):       # each call appends one more measurement per landmark (histories grow by 20)
        r = lmap.update_frame(last, h["world_to_camera"][last], h["camera_to_world"][last], ids, cam)
    ms = (time.perf_counter() - t0) / reps * 1e3
    out.append({"landmarks": n, "measurements": int(lengths.sum()) + n * (reps // 2), "ms_per_call": ms, "resident": True,
                "what": "vslam_landmark_map_update_frame: ids + camera coordinates in (28 KB), world / updates / outcome "
                        "out (29 KB), histories resident in HBM; python harness",
                "mean_iterations": float(r["iterations"].mean()), "adopted_fraction": float((r["outcome"] == 1).mean())})
    lmap.close()
    return out


def sequence_latency(api, configs, synth, device):
    """BASELINE.json configs[0]/[1] shape: ONE sequence, frame by frame through the reference-shaped calls
    (initialize -> compute, thresholds fed back between frames, host images in, host framepoints out, one host
    synchronisation per call): single-stream frames/s, i.e. the latency-bound regime of a live tracker."""
    out = {}
    for name in ("kitti", "euroc", "hd"):     # hd: BASELINE.json configs[4] shape (1920x1080, 4k bins), one sequence
        cfg = configs.BY_NAME[name]
        cam = synth.camera(cfg.camera)
        world = synth.BandWorld(cam.cols, cam.rows, 7, max_frames=40)
        frames = []
        for k in range(24):       # page-locked frame buffers (vslam_host_alloc): the H2D copy is one async DMA
            l, r = world.pair(k)
            pl, pr = api.pinned_empty(l.shape), api.pinned_empty(r.shape)
            pl[:], pr[:] = l, r
            frames.append((pl, pr))
        gen = api.StereoFramePointGenerator(cfg, cam, device=device)
        for k in range(4):
            gen.initialize(frames[k][0], frames[k][1], k == 0)
            gen.compute()
        t0 = time.perf_counter()
        n_fp = 0
        for k in range(4, 24):
            gen.initialize(frames[k][0], frames[k][1], False)
            n_fp += len(gen.compute())
        dt = time.perf_counter() - t0
        out[name] = {"frames_per_s": 20 / dt, "ms_per_frame": dt / 20 * 1e3, "mean_framepoints": n_fp / 20,
                     "image": "%dx%d" % (cam.cols, cam.rows)}
        # the tracker's per-frame order (pose_tracker_3d.cpp:80, 239, 210): initialize -> track against ALL points of the
        # previous frame -> compute with the tracks pre-loaded from device memory.  The host part (descriptor download,
        # assembly of the previous points) is inside the timed region.
        T = np.hstack([np.eye(3), np.zeros((3, 1))])
        T[0, 3] = -(-cam.bx / cam.fx) / 4
        prev, n_tracks, n_prev, n_new = None, 0, 0, 0
        t_total = 0.0
        for k in range(24):
            t0 = time.perf_counter()
            gen.initialize(frames[k][0], frames[k][1], k == 0)
            parts = []
            if prev is not None:
                r = gen.track(prev, T, False, 15, 25.6)
                parts.append(r["tracks"])
                new = gen.compute(api.TRACKED_FROM_LAST_TRACK)
            else:
                new = gen.compute()
            parts.append(new)
            _, dl = gen.features(0)
            _, dr = gen.features(1)
            nxt = api.make_previous_points(parts, dl, dr)
            t1 = time.perf_counter()
            if k >= 4:
                t_total += t1 - t0
                n_tracks += len(parts[0]) if prev is not None else 0
                n_prev += len(prev)
                n_new += len(new)
            prev = nxt
        out[name]["tracked"] = {"frames_per_s": 20 / t_total, "ms_per_frame": t_total / 20 * 1e3,
                                "mean_previous_points": n_prev / 20, "mean_tracks": n_tracks / 20,
                                "mean_new_points": n_new / 20, "host": "python harness"}
        gen.close()
        # the same loop from C++14 (tools/sequence_runner.cpp over include/vslam_b200.hpp): what a native host pays
        out[name]["tracked_native"] = native_sequence(cfg, cam, frames)
        # ... and the frame as one device pass (vslam_fpg_frame_step)
        out[name]["tracked_fused"] = native_sequence(cfg, cam, frames, fused=True)
        # ... with the next frame's images uploaded while the current frame runs (vslam_fpg_frame_step_prefetch)
        out[name]["tracked_fused_prefetch"] = native_sequence(cfg, cam, frames, fused=2)
    return out


def sequence_multi(configs, synth, rank, world, device, barrier, dist, torch, dev, frames_n=100, passes=5):
    """BASELINE.json configs[4]: 8 independent 1920x1080 sequences (seeds 8000 + rank, 100 frames, 84 x 47 = 3948 bins),
    one per GPU: every rank drives its own sequence through initialize -> track -> StereoUVAligner initialize + converge
    -> errors / inliers -> compute from C++14 (tools/sequence_runner.cpp).  Rank 0 first runs ALONE, then all ranks run
    at once: efficiency = mean concurrent frames/s / rank 0's frames/s alone."""
    cfg = configs.BY_NAME["hd"]
    cam = synth.camera(cfg.camera)
    bw = synth.BandWorld(cam.cols, cam.rows, 8000 + rank, max_frames=frames_n)
    frames = [bw.pair(k) for k in range(frames_n)]
    alone = native_sequence(cfg, cam, frames, passes=passes, device=device) if rank == 0 else None
    alone_fused = native_sequence(cfg, cam, frames, passes=passes, device=device, fused=True) if rank == 0 else None
    barrier()
    mine = native_sequence(cfg, cam, frames, passes=passes, device=device)
    barrier()
    mine_fused = native_sequence(cfg, cam, frames, passes=passes, device=device, fused=True)
    barrier()
    alone_prefetch = native_sequence(cfg, cam, frames, passes=passes, device=device, fused=2) if rank == 0 else None
    barrier()
    mine_prefetch = native_sequence(cfg, cam, frames, passes=passes, device=device, fused=2)
    barrier()
    fps = float(mine["frames_per_s"]) if mine else 0.0
    fps_fused = float(mine_fused["frames_per_s"]) if mine_fused else 0.0
    fps_prefetch = float(mine_prefetch["frames_per_s"]) if mine_prefetch else 0.0
    if dist is not None:
        t = torch.zeros(3 * world, dtype=torch.float64, device=dev)
        t[rank] = fps
        t[world + rank] = fps_fused
        t[2 * world + rank] = fps_prefetch
        dist.all_reduce(t)
        per_rank = [float(x) for x in t.tolist()[:world]]
        per_rank_fused = [float(x) for x in t.tolist()[world:2 * world]]
        per_rank_prefetch = [float(x) for x in t.tolist()[2 * world:]]
    else:
        per_rank, per_rank_fused, per_rank_prefetch = [fps], [fps_fused], [fps_prefetch]
    if rank != 0:
        return None
    out = {"workload": "%d-frame 1920x1080 band-world sequence per GPU (seeds 8000 + rank), bin 23, FAST thr 20..100; "
                       "per frame initialize + track + StereoUVAligner converge + compute; sequence replayed %d times"
                       % (frames_n, passes),
           "per_rank_frames_per_s": per_rank, "aggregate_frames_per_s": sum(per_rank),
           "rank0_alone": alone, "rank0_concurrent": mine,
           "efficiency": (sum(per_rank) / len(per_rank)) / alone["frames_per_s"] if alone else None,
           "fused": {"per_rank_frames_per_s": per_rank_fused, "aggregate_frames_per_s": sum(per_rank_fused),
                     "rank0_alone": alone_fused, "rank0_concurrent": mine_fused,
                     "efficiency": (sum(per_rank_fused) / len(per_rank_fused)) / alone_fused["frames_per_s"]
                     if alone_fused else None},
           "fused_prefetch": {"what": "fused frames, the images of frame k + 1 uploaded while frame k runs "
                                      "(vslam_fpg_frame_step_prefetch)",
                              "per_rank_frames_per_s": per_rank_prefetch,
                              "aggregate_frames_per_s": sum(per_rank_prefetch),
                              "rank0_alone": alone_prefetch, "rank0_concurrent": mine_prefetch,
                              "efficiency": (sum(per_rank_prefetch) / len(per_rank_prefetch))
                              / alone_prefetch["frames_per_s"] if alone_prefetch else None}}
    return out


_RUNNER = {}


def native_sequence(cfg, cam, frames, warmup=4, acfg=None, passes=1, device=0, fused=False):
    """builds tools/sequence_runner.cpp once (g++, C++14) and runs the tracked sequence through it; None when the
    compiler is missing or anything fails -- the number is informative, the run stays valid without it"""
    import tempfile
    try:
        if "exe" not in _RUNNER:
            import atexit
            import shutil
            d = tempfile.mkdtemp(prefix="vslam_runner_")
            atexit.register(shutil.rmtree, d, True)
            exe = os.path.join(d, "sequence_runner")
            pkg = os.path.join(ROOT, "vslam-pose-estimation-framework_b200")
            subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"),
                                   os.path.join(ROOT, "tools", "sequence_runner.cpp"), "-o", exe, "-L", pkg, "-lvslam_b200",
                                   "-Wl,-rpath," + pkg], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            _RUNNER["exe"], _RUNNER["dir"] = exe, d
        # the frames file is written once per sequence and reused by the stepwise / fused / prefetch runs of it
        key = (cfg.name, len(frames), id(frames))
        path = _RUNNER.setdefault("frames", {}).get(key)
        if path is None:
            for old in _RUNNER["frames"].values():
                if os.path.exists(old):
                    os.remove(old)
            _RUNNER["frames"].clear()
            path = os.path.join(_RUNNER["dir"], "frames_%s_%d.u8" % (cfg.name, len(_RUNNER["frames"])))
            with open(path, "wb") as f:
                for l, r in frames:
                    f.write(np.ascontiguousarray(l).tobytes())
                    f.write(np.ascontiguousarray(r).tobytes())
            _RUNNER["frames"][key] = path
        args = [cam.rows, cam.cols, cfg.target_number_of_keypoints_tolerance, cfg.detector_threshold_minimum,
                cfg.detector_threshold_maximum, cfg.detector_threshold_maximum_change, cfg.number_of_detectors_vertical,
                cfg.number_of_detectors_horizontal, int(cfg.enable_keypoint_binning), cfg.bin_size_pixels,
                cfg.maximum_matching_distance_triangulation, cfg.minimum_disparity_pixels,
                cfg.maximum_epipolar_search_offset_pixels, cam.fx, cam.fy, cam.cx, cam.cy, cam.bx, 15]
        if acfg is None:
            from vslam_b200 import configs as _c
            acfg = _c.ALIGNER_BY_NAME[cfg.name]
        args += [acfg.error_delta_for_convergence, acfg.maximum_error_kernel, acfg.damping,
                 acfg.minimum_number_of_inliers, acfg.maximum_reliable_depth_meters]
        env = dict(os.environ)
        if device:      # the runner uses device 0 of what it sees: show it this rank's GPU only
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            ids = visible.split(",") if visible else [str(i) for i in range(64)]
            env["CUDA_VISIBLE_DEVICES"] = ids[device]
        res = subprocess.run([_RUNNER["exe"], path, str(len(frames)), str(warmup)] + [repr(a) for a in args]
                             + [str(passes), str(int(fused))], capture_output=True, text=True, timeout=300, env=env)
        if res.returncode != 0:
            return None
        r = json.loads(res.stdout.strip().splitlines()[-1])
        for line in res.stderr.splitlines():      # VSLAM_RUNNER_PROFILE=1: device us per stage of the fused frame
            if line.startswith("device us per frame"):
                r["stage_profile"] = line
        r["host"] = "C++14 (tools/sequence_runner.cpp)"
        if fused:
            r["what"] = ("vslam_fpg_frame_step: the frame as ONE graph launch and one synchronisation, points() of the "
                         "previous frame resident on the device; results bit-identical to the stepwise calls")
            for key in list(r):
                if key.startswith("us_"):
                    del r[key]
        return r
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


_RESULT = sys.stdout


def _claim_stdout():
    """stdout carries ONE JSON line: keep a private handle on the real stdout for it and point file descriptor 1 at
    stderr, so that nothing a library prints (NCCL's version banner under torchrun, for one) can land beside it"""
    global _RESULT
    sys.stdout.flush()
    _RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from vslam_b200 import api, configs, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # (and file descriptor 1 is stderr by now, _claim_stdout)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else "single rank: not bound"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        return sharding.max_over_ranks(x, dev)

    cfg, acfg = configs.BY_NAME[CONFIG_NAME], configs.ALIGNER_BY_NAME[CONFIG_NAME]
    cam = synth.camera(cfg.camera)
    R, K, W = args.rounds, args.steps, max(args.warmup, 0)
    cores = len(os.sched_getaffinity(0))
    if args.scaling == "strong":
        # BASELINE.json configs[2] as written: args.pairs pairs in TOTAL, pair i (seed i) on rank floor(i * G / B)
        part = sharding.block_partition(args.pairs, world, rank)
        lo, hi = part.start, part.stop
        P = hi - lo
        D = min(args.distinct or P, P)
        seeds = range(lo, lo + D)
        P_total = args.pairs
    else:
        P = args.pairs
        D = min(args.distinct or 256, P)
        seeds = sharding.weak_seeds(D, rank)
        P_total = world * P

    # ---- synthetic inputs: D distinct band-world pairs per rank (seeds disjoint across ranks), tiled to P pairs in
    # pinned host memory (every pair is its own memory and is processed independently)
    dl, dr = synth.band_world_batch(cfg.camera, seeds, workers=max(1, min(cores // world, 32)))
    if args.wc_images:      # write-combined upload sources (the host only writes them): DMA reads skip the cache snoop
        os.environ["VSLAM_HOST_ALLOC_WC"] = "1"
    left = api.pinned_empty((P, cam.rows, cam.cols))
    right = api.pinned_empty((P, cam.rows, cam.cols))
    os.environ.pop("VSLAM_HOST_ALLOC_WC", None)
    for i in range(0, P, D):
        n = min(D, P - i)
        left[i:i + n], right[i:i + n] = dl[:n], dr[:n]

    gen = api.StereoFramePointGenerator(cfg, cam, device=local_rank, max_batch=P)
    out = api.pinned_empty((P, gen.out_capacity), api.FRAMEPOINT)
    counts = np.zeros(P, np.int32)
    T = _prior()
    stream = torch.cuda.ExternalStream(gen.stream, device=dev)

    def step_resident():
        gen.batch_run(P, True)
        gen.batch_linearize(P, T, acfg, False, R)

    def step_e2e():
        gen.batch_process(left, right, True, out, counts)
        gen.batch_linearize(P, T, acfg, False, R)
        return gen.batch_systems(P, raw=True)

    # ---- device-resident throughput
    gen.batch_upload(left, right)
    with ClockSampler(local_rank) as clk:
        for _ in range(W):
            step_resident()
        barrier()
        launches0 = gen.launch_count
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk.begin()
        start.record(stream)
        for _ in range(K):
            step_resident()
        end.record(stream)
        barrier()
        clk.end()
    ms = max_over_ranks(start.elapsed_time(end))
    launches = gen.launch_count - launches0
    value = P_total * K / (ms * 1e-3)

    # ---- end to end through the host-buffer API
    for _ in range(min(W, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = P_total * K / e2e_s
    # copy-only control: the same pinned buffers through the same chunked pipeline with the kernels left out (only the
    # re-pitch kernel runs) -- if this is as slow as e2e, the host -> device path is the limit, not the pipeline
    def step_copies_only():
        gen.batch_upload(left, right)          # host -> device: the images
        gen.batch_download(P, out)             # device -> host: the framepoint records (of the last run) ...
        return gen.batch_systems(P, raw=True)  # ... and the normal equations
    step_copies_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_copies_only()
    barrier()
    h2d_s = max_over_ranks(time.perf_counter() - t0)
    h2d = 2 * P * cam.rows * cam.cols
    d2h = P * gen.out_capacity * api.FRAMEPOINT.itemsize + P * 8 * 4 + P * 32 * 8

    # ---- per-kernel durations (CUDA events inside the library, serialised on one stream) and the roofline
    _, nf, nm, nl, nr = gen.batch_download(P, out)
    gen.set_profiling(-1)
    gen.set_profiling(1)
    for _ in range(max(1, min(K, 3))):
        step_resident()
    gen.synchronize()
    prof = gen.kernel_profile()
    gen.set_profiling(0)
    steps_prof = max(1, min(K, 3))
    chunks = prof[KERNEL_OF_INTEREST][1] / steps_prof
    kernel_ms = {k: v[0] / steps_prof for k, v in prof.items()}
    total_kernel_ms = sum(kernel_ms.values())
    peak, peak_kind = peaks()
    # algorithmic bytes of the FAST kernel per step: both images read once (u8) + 8 B per raw keypoint found
    # (2*W*H + 2*N_kp*8 of the per-frame figure); N_kp from this run's own counts is not kept per image,
    # the descriptor-valid count (nl + nr) is a lower bound and is what is charged
    fast_bytes = float(2 * P * cam.rows * cam.cols + 8 * (nl.sum() + nr.sum()))
    fast_ms_per_launch = kernel_ms[KERNEL_OF_INTEREST] / max(chunks, 1)
    achieved = fast_bytes / max(chunks, 1) / (fast_ms_per_launch * 1e-3) / 1e9
    # whole path, per frame: 2WH + 2 N_kp 8 + 2 N_desc 32 + 2 N_desc 40 + N_match 48  (+ 81 B per point per round)
    n_desc = float(nl.sum() + nr.sum())
    path_bytes = (2.0 * P * cam.rows * cam.cols + 8 * n_desc + 32 * n_desc + 40 * n_desc + 48.0 * nm.sum()
                  + 81.0 * nf.sum() * R)
    # DRAM traffic of the kernel from the committed ncu --set full capture, scaled to this run's images per launch
    traffic, limiters = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["fast_nms_kernel"]
        if t["image"] == "%dx%d u8" % (cam.cols, cam.rows):
            traffic = t["dram_bytes_per_launch"] / t["images_per_launch"] * (2.0 * P / max(chunks, 1))
        limiters = t.get("limiters_pct_of_peak")     # from the same committed capture: what really bounds the kernel
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "fast_nms_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": fast_bytes / max(chunks, 1),
                "peak_source": peak_kind + " (of %s)" % peak_kind,
                "launch_ms": fast_ms_per_launch, "launches_per_step": chunks,
                "kernel_share_of_step": kernel_ms[KERNEL_OF_INTEREST] / total_kernel_ms,
                "kernel_ms_per_step": kernel_ms, "ncu_limiters_pct_of_peak": limiters,
                "path": {"algorithmic_bytes_per_step": path_bytes,
                         "achieved_gbs": path_bytes / (ms / K * 1e-3) / 1e9,
                         "frac": path_bytes / (ms / K * 1e-3) / 1e9 / peak}}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8+f64", "data": "synthetic (band-world, %d distinct pairs per GPU tiled to %d)" % (D, P),
            "config": workload_config(args, args.pairs), "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / K * 1e3, "host_placement": numa,
                    "image_buffers": "write-combined pinned" if args.wc_images else "pinned",
                    "h2d_gbs_per_gpu": h2d * K / e2e_s / 1e9,
                    "copy_only_control": {"ms_per_step": h2d_s / K * 1e3, "h2d_only_gbs_per_gpu": h2d * K / h2d_s / 1e9,
                                          "h2d_only_gbs_all_gpus": world * h2d * K / h2d_s / 1e9,
                                          "frames_per_s_if_copy_bound": P_total * K / h2d_s,
                                          "e2e_over_copy_only": h2d_s / e2e_s,
                                          "what": "the step's copies without its kernels: vslam_fpg_batch_upload of the "
                                                  "same pinned image buffers (same chunks and lanes; only the row re-pitch "
                                                  "kernel runs), then the download of the framepoint records and of the "
                                                  "normal equations into the same host buffers; max over ranks"}},
            "gpu_launches": int(launches), "roofline": roofline,
            "counts": {"mean_descriptors_left": float(nl.mean()), "mean_matches": float(nm.mean()),
                       "mean_framepoints": float(nf.mean())}}
    if not args.no_extras:
        # BASELINE.json configs[4]: one independent 1920x1080 sequence per GPU, every rank runs its own
        multi = sequence_multi(configs, synth, rank, world, local_rank, barrier, dist if world > 1 else None, torch, dev)
        if rank == 0:
            line["sequence_multi"] = multi
    if rank == 0 and not args.no_extras:
        line["aligner_stress"] = aligner_stress(api, configs, synth, torch, dev)
        line["sequence"] = sequence_latency(api, configs, synth, local_rank)
        line["landmark_refinement"] = landmark_refinement(api, synth, local_rank)
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(dl, dr, R, args.cpu_sample)
        if not args.no_extras:
            # BASELINE.json configs[0]: the 200-frame KITTI sequence through the reference's tracker, CPU classes and
            # GPU adapters side by side (same unmodified PoseTracker3D, same frames)
            cam0 = synth.camera("kitti")
            world0 = synth.BandWorld(cam0.cols, cam0.rows, 1, max_frames=200)
            frames0 = [world0.pair(k) for k in range(200)]
            seq_cpu = cpu_sequence_baseline(frames0)
            line["cpu_baseline"]["sequence"] = seq_cpu
            seq_gpu = gpu_sequence_through_the_reference_tracker(frames0)
            line.setdefault("sequence", {})["reference_tracker_kitti_200"] = {
                "gpu_adapters": seq_gpu, "cpu_reference": seq_cpu,
                "speedup": (seq_gpu["value"] / seq_cpu["value"]) if seq_gpu and seq_cpu else None}
    if rank == 0:
        print(json.dumps(line), file=_RESULT, flush=True)
    gen.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
