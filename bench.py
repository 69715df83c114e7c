#!/usr/bin/env python
"""bench.py -- frames/s of ORB + LSD/LBD extraction (BASELINE.json metric) on N B200, and Hamming matches/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5]
  (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Configurations = BASELINE.json `configs`:
  c4 (default, the headline): batched sequence extraction, 1920x1080 frames, 2000 ORB (8 lv x1.2, FAST 20/7) + LSD/LBD 800 lines
      (TUM LSD options, min length 0.02 * 1080), 1024 frames per GPU per step in sub-batches, frame i -> rank i mod N.
  c2: EuRoC-style stereo 752x480 pairs, 1200 ORB + 200 lines per image (round 1's headline).
  c3: KITTI-style stereo 1241x376 pairs, 2000 ORB + 800 lines, plus stereo point matching and L<->R line matchNNR.
  c1: one 640x480 frame, ORB 1000 + TUM lines through the drop-in calls (ORB thread || line thread): latency.
  c5: brute-force Hamming top-2 sweep, 1e4 queries x 1e4..1e7 train rows, train-sharded with an NCCL merge.
A default run prints the c4 line and attaches short measurements of the others under `other_configs`.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM (device-timed, max over ranks); `e2e` = the
same through the C-ABI host-buffer calls (pinned host memory, H2D + D2H inside the timed region); `roofline` = the kernel
that takes the largest share of the step measured while all streams run together, against the measured HBM peak, plus the
whole step's algorithmic bytes / step time; `cpu_baseline` = the reference's own CPU code (oracle/_ref, compiled from the
reference's unmodified sources) on the box's host cores on a bounded sample.
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
sys.path.insert(0, ROOT)

ORB_TUM = dict(nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)
LSD_TUM = dict(nlevels=2, refine=0, scale=1.1, sigma_scale=0.6, quant=2.2, ang_th=12.5, log_eps=1.0, density_th=0.6, n_bins=1024)
CONFIGS = {
    # Examples/Monocular/TUM1.yaml: ORB 1000, Lineextractor 600 lines, min_line_length_ratio 0.02 (src/Tracking.cc:168-169)
    "c1": dict(W=640, H=480, stereo=False, frames=2, sub=1, orb=dict(ORB_TUM),
               line=dict(LSD_TUM, nfeatures=600, min_line_length=0.02 * 480),
               workload="single 640x480 frame, ORB 1000 (8 lv x1.2, FAST 20/7) + LSD/LBD 600 lines (TUM settings), drop-in calls"),
    # Examples/Monocular/EuRoC.yaml line settings (Camera.width absent -> min length 0)
    "c2": dict(W=752, H=480, stereo=True, frames=1024, sub=512, orb=dict(ORB_TUM, nfeatures=1200),
               line=dict(LSD_TUM, nfeatures=200, sigma_scale=0.8, density_th=0.8, min_line_length=0.0),
               workload="EuRoC-style stereo 752x480 pairs, 1200 ORB (8 lv x1.2, FAST 20/7) + LSD/LBD 200 lines per image"),
    "c3": dict(W=1241, H=376, stereo=True, frames=512, sub=256, orb=dict(ORB_TUM, nfeatures=2000),
               line=dict(LSD_TUM, nfeatures=800, min_line_length=0.02 * 376),
               workload="KITTI-style stereo 1241x376 pairs, 2000 ORB + LSD/LBD 800 lines per image, stereo point matching + L<->R line matchNNR"),
    "c4": dict(W=1920, H=1080, stereo=False, frames=1024, sub=512, orb=dict(ORB_TUM, nfeatures=2000),
               line=dict(LSD_TUM, nfeatures=800, min_line_length=0.02 * 1080),
               workload="batched sequence extraction: 1920x1080 frames, 2000 ORB (8 lv x1.2, FAST 20/7) + LSD/LBD 800 lines (TUM LSD options), frame-sharded"),
}
LINE_CONTEXTS = 2             # line extractor instances (own context + host thread each; every instance runs its two octaves on two streams)
ORB_CONTEXTS = 2              # ORB extractor instances, same idea (uploads of one overlap kernels of the other)
NSETS = 2                     # distinct frame sets per GPU; step k works on set k % NSETS


def _gauss_kernel_q8(ksize, sigma):
    """cv::GaussianBlur's 8-bit kernel: Q8, error-diffused, sums to 256."""
    import math
    n2 = ksize // 2
    scale2x = -0.125 / (sigma * sigma)
    w = [math.exp(float(x * x) * scale2x) for x in range(1 - ksize, 0, 2)]
    tot = 1.0 / (sum(w) * 2 + 1)
    q, err, acc = [0] * ksize, 0.0, 0
    for i in range(n2):
        v = w[i] * tot * 256.0 + err
        v0 = int(np.rint(v))
        err = v - v0
        q[i] = q[ksize - 1 - i] = v0
        acc += v0
    q[n2] = 256 - 2 * acc
    return q


def _gauss_blur_u8(img, ksize, sigma):
    """Separable Q8 Gaussian with BORDER_REFLECT_101, (v + 32768) >> 16 -- integer arithmetic, bit-identical to
    cv::GaussianBlur on 8U (scipy's correlate1d in int32 with mode 'mirror' = reflect-101)."""
    from scipy.ndimage import correlate1d
    q = np.array(_gauss_kernel_q8(ksize, sigma), np.int32)
    h = correlate1d(img.astype(np.int32), q, axis=1, mode="mirror")
    v = correlate1d(h, q, axis=0, mode="mirror")
    return ((v + 32768) >> 16).astype(np.uint8)


def synth_image(w, h, seed):
    """The synthetic test image of SURVEY.md 8d (uniform u8 noise -> Gaussian s=2 -> K filled rectangles -> 3x3 s=0.8 blur).
    Input generation only -- numpy / scipy, nothing of oracle/ is involved (tests/test_bench_inputs.py checks that it equals
    the generator the parity tests use)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    img = _gauss_blur_u8(img, 13, 2.0)
    K = int(round(60.0 * (w * h) / (640.0 * 480.0)))
    for _ in range(K):
        x0 = int(rng.integers(0, w - 8)); y0 = int(rng.integers(0, h - 8))
        rw = int(rng.integers(8, max(9, w // 4))); rh = int(rng.integers(8, max(9, h // 4)))
        g = int(rng.integers(0, 256))
        img[y0:min(h, y0 + rh), x0:min(w, x0 + rw)] = g
    return _gauss_blur_u8(img, 3, 0.8)


def make_frames(cfg, ids):
    """Frames with the given global ids.  Mono: frame i = synth_image(seed i).  Stereo: frames 2p, 2p+1 = left, right of pair p
    (right = left shifted by a disparity + small noise)."""
    from concurrent.futures import ThreadPoolExecutor
    W, H = cfg["W"], cfg["H"]

    def one(i):
        if not cfg["stereo"]:
            return synth_image(W, H, i)
        p = i // 2
        left = synth_image(W, H, p)
        if i % 2 == 0:
            return left
        rng = np.random.default_rng(10_000 + p)
        right = np.roll(left, -int(rng.integers(4, 24)), axis=1).astype(np.int16) + rng.integers(-2, 3, left.shape)
        return np.clip(right, 0, 255).astype(np.uint8)

    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        return np.stack(list(ex.map(one, ids)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference legs
def cpu_reference_throughput(cfg, frames, nthreads, what="both"):
    """The reference's own CPU code on `frames`, one frame per worker thread: ORBextractor::operator() +
    Lineextractor::ComputeLsdWithLbd compiled from the reference's unmodified sources (oracle/_ref, kind "reference");
    falls back to the C restatement (oracle/, kind "port") only if libref.so is absent.  Returns (frames/s, seconds, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O, ref as R
    O.build()
    kind = "reference" if (R.available() or R.build()) else "port"
    M = R if kind == "reference" else O
    if kind == "reference":
        R.set_heap_mode(0)          # the reference as it runs (glibc malloc)
    prm = O.line_params(**cfg["line"])
    orb = cfg["orb"]
    local = threading.local()

    def work(i):
        n = 0
        if what in ("both", "orb"):
            if not hasattr(local, "orb"):
                local.orb = M.ORBextractor(orb["nfeatures"], orb["scaleFactor"], orb["nlevels"], orb["iniThFAST"], orb["minThFAST"])
            n += len(local.orb(frames[i])[0])
        if what in ("both", "line"):
            n += len(M.line_extract(prm, frames[i])[0])
        return n

    with ThreadPoolExecutor(nthreads) as ex:
        list(ex.map(work, range(min(len(frames), nthreads))))   # warm-up (ctypes releases the GIL)
        t0 = time.perf_counter()
        list(ex.map(work, range(len(frames))))
        dt = time.perf_counter() - t0
    return len(frames) / dt, dt, kind


def cpu_knn2(q, t, nthreads):
    """Brute-force Hamming top-2 on the host: cv2.BFMatcher.knnMatch(k=2) -- what Linematcher::matchNNR calls
    (src/Linematcher.cc:526-527) -- with `nthreads` OpenCV threads; seconds."""
    import cv2
    cv2.setNumThreads(nthreads)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    t0 = time.perf_counter()
    bf.knnMatch(q, t, 2)
    return time.perf_counter() - t0


def config_dict(name, cfg, B, world, extra=None):
    d = {"workload": cfg["workload"], "config": name, "frames_per_gpu_per_step": B, "frame": "one %dx%d image" % (cfg["W"], cfg["H"]) +
         ("; a stereo pair is 2 frames" if cfg["stereo"] else ""),
         "sharding": "frame i -> rank i mod N" + (" (left/right of a pair on separate GPUs for N > 1)" if cfg["stereo"] else "") + ", no collective"}
    if extra:
        d.update(extra)
    return d


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, all threads, each step a
    bounded sample of the same workload (same config dict as our arm)."""
    if rank != 0:
        return
    name = args.config
    cfg = CONFIGS[name if name != "c5" else "c4"]
    cores = os.cpu_count() or 1
    if name == "c5":
        return run_reference_c5(args, cores)
    B = args.frames or cfg["frames"]
    # sample sized for ~10-20 s of CPU work per step
    per_frame_s = 1.1e-6 * cfg["W"] * cfg["H"] * 0.25      # ~0.5 s at 1080p, ~0.09 s at 752x480 (one thread, -O3)
    nfr = max(cores, min(B, int(12.0 * cores / per_frame_s)))
    nfr -= nfr % 2
    frames = make_frames(cfg, list(range(nfr)))
    kind = "reference"
    for _ in range(1 if args.warmup else 0):
        cpu_reference_throughput(cfg, frames[:cores], cores)
    tot_t, tot_f = 0.0, 0
    for _ in range(args.steps):
        fps, dt, kind = cpu_reference_throughput(cfg, frames, cores)
        tot_t += dt; tot_f += len(frames)
    value = tot_f / tot_t
    line = {"impl": "reference", "metric": "frames/s ORB+LSD/LBD extraction", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(name, cfg, B, world),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "%d of the step's %d frames per step, one frame per thread on %d threads (%s)" %
                                       (len(frames), B, cores, "oracle/_ref: the reference's own sources, -O3 -march=native" if kind == "reference" else "C oracle, -O3")},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_c5(args, cores):
    NQ, NT = 10000, 1000000
    rng = np.random.default_rng(1234)
    q = rng.integers(0, 256, (NQ, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (NT // 10, 32), dtype=np.uint8)     # bounded sample: a tenth of the train rows, same queries
    tot = 0.0
    for _ in range(args.steps):
        tot += cpu_knn2(q, t, cores)
    pairs = NQ * (NT // 10) * args.steps / tot
    value = pairs / NT            # queries/s at 1e6 train rows
    line = {"impl": "reference", "metric": "Hamming matches/s", "value": value, "unit": "queries/s at 1e6 train rows", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": c5_config(1, NT),
            "cpu_baseline": {"value": value, "unit": "queries/s at 1e6 train rows", "cores": cores, "kind": "reference",
                             "sample": "cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2), 1e4 queries x 1e5 train rows per step on %d threads, scaled by pairs" % cores},
            "e2e": {"value": value, "unit": "queries/s at 1e6 train rows", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def c5_config(world, nt):
    return {"workload": "brute-force Hamming top-2 + ratio 0.75, 256-bit descriptors, 1e4 queries x %.0e train rows" % nt, "config": "c5",
            "sharding": "train rows sharded over %d GPU(s)%s" % (world, ", NCCL all-gather of the per-shard top-2 + merge" if world > 1 else "")}


# ------------------------------------------------------------------------------------------------ GPU arm
STAGGER_MS = float(os.environ.get("PLF_BENCH_STAGGER_MS", "0"))


class Extraction:
    """One configuration's extraction workload on this rank's GPU: device-resident and end-to-end runs."""

    def __init__(self, S, torch, cfg, B, rank, world, dev, args, nsets=NSETS):
        self.S, self.torch, self.cfg, self.B, self.dev = S, torch, cfg, B, dev
        W, H = cfg["W"], cfg["H"]
        self.W, self.H = W, H
        orb, line = cfg["orb"], cfg["line"]
        # global frame i -> rank i mod world; set s holds global frames [s * B * world, (s + 1) * B * world)
        self.nsets = nsets
        sets = []
        for s in range(nsets):
            ids = [s * B * world + i for i in range(B * world) if i % world == rank]
            sets.append(make_frames(cfg, ids))
        self.frames = sets
        NO = args.orb_contexts if B % (2 * args.orb_contexts) == 0 else 1
        NL = args.line_contexts if B % (2 * args.line_contexts) == 0 else 1
        self.NO, self.NL = NO, NL
        self.ctx_os = [S.Context(dev, priority=-1) for _ in range(NO)]    # ORB streams: filler priority, run one step ahead of the lines
        self.ctx_ls = [S.Context(dev, priority=1) for _ in range(NL)]     # line streams: their region-growing chains set the step time
        self.ctx_o = self.ctx_os[0]
        self.lib = self.ctx_o.lib
        self.orbs = [S.ORBextractor(orb["nfeatures"], orb["scaleFactor"], orb["nlevels"], orb["iniThFAST"], orb["minThFAST"], ctx=c) for c in self.ctx_os]
        self.les = [S.Lineextractor(line["nfeatures"], line["nlevels"], line["refine"], line["scale"], line["sigma_scale"], line["quant"],
                                    line["ang_th"], line["log_eps"], line["density_th"], line["n_bins"], line["min_line_length"], ctx=c) for c in self.ctx_ls]
        self.capk, self.capl = self.orbs[0].max_keypoints, self.les[0].max_keylines
        capk, capl = self.capk, self.capl
        self.BL, self.BO = B // NL, B // NO
        sub = args.line_sub or cfg["sub"]
        self.SUB = sub if sub and self.BL % sub == 0 else self.BL
        osub = args.orb_sub or min(self.BO, 256)
        self.OSUB = osub if self.BO % osub == 0 else self.BO
        from concurrent.futures import ThreadPoolExecutor
        self.pool = ThreadPoolExecutor(NL + NO + 1)
        self.d_img = [torch.from_numpy(f).cuda() for f in sets]
        self.d_kps = torch.empty((B, capk, 28), dtype=torch.uint8, device="cuda"); self.d_desc = torch.empty((B, capk, 32), dtype=torch.uint8, device="cuda")
        self.d_nk = torch.empty(B, dtype=torch.int32, device="cuda")
        self.d_kl = torch.empty((B, capl, 68), dtype=torch.uint8, device="cuda"); self.d_mid = torch.empty((B, capl, 28), dtype=torch.uint8, device="cuda")
        self.d_ld = torch.empty((B, capl, 32), dtype=torch.uint8, device="cuda"); self.d_nl = torch.empty(B, dtype=torch.int32, device="cuda")
        self.h_img = None

    def close(self):
        self.pool.shutdown()
        for o in self.orbs + self.les:
            o.close()
        for c in self.ctx_os + self.ctx_ls:
            c.close()
        self.d_img = self.h_img = None
        self.torch.cuda.empty_cache()

    def contexts(self):
        return self.ctx_os + self.ctx_ls

    # ---- device-resident ----
    def dev_orb_i(self, i, k):
        W, H = self.W, self.H
        img = self.d_img[k % self.nsets]
        for j in range(i * self.BO, (i + 1) * self.BO, self.OSUB):
            s = slice(j, j + self.OSUB)
            self.ctx_os[i].check(self.lib.plf_orb_extract_batch_device(self.orbs[i].h, img[s].data_ptr(), self.OSUB, W, H, W, W * H, self.d_kps[s].data_ptr(),
                                                                       self.d_desc[s].data_ptr(), self.capk, self.d_nk[s].data_ptr()))

    def dev_orb(self, k):
        for i in range(self.NO):   # asynchronous launches, one stream per ORB context
            self.dev_orb_i(i, k)

    def dev_line(self, i, k):
        W, H = self.W, self.H
        img = self.d_img[k % self.nsets]
        for j in range(i * self.BL, (i + 1) * self.BL, self.SUB):     # sub-batches bound the workspace and keep the contexts out of lockstep
            s = slice(j, j + self.SUB)
            self.ctx_ls[i].check(self.lib.plf_line_extract_batch_device(self.les[i].h, img[s].data_ptr(), self.SUB, W, H, W, W * H, self.d_kl[s].data_ptr(),
                                                                        self.d_mid[s].data_ptr(), self.d_ld[s].data_ptr(), self.capl, self.d_nl[s].data_ptr()))

    def _line_steps(self, i, nsteps, with_orb):
        if STAGGER_MS and i:
            time.sleep(i * STAGGER_MS * 1e-3)    # contexts out of phase: one grows regions (latency chain) while the other streams
        for k in range(nsteps):
            if with_orb and i == 0 and k + 1 < nsteps:
                self.dev_orb(k + 1)      # ORB launches of step k + 1 (low-priority streams): filler for the gaps the line path leaves
            self.dev_line(i, k)

    def run_device(self, nsteps, what="both"):
        """nsteps steps back to back: every extractor instance (own stream + host thread) walks through its share of each
        step's batch without a global barrier between steps, so the tail of one step overlaps the start of the next, as it
        does in a running system.  Returns the device time in ms."""
        self.torch.cuda.synchronize()
        self.ctx_o.timer_start()
        if what == "orb":
            for k in range(nsteps):
                self.dev_orb(k)
        else:
            if what == "both":
                self.dev_orb(0)
            futs = [self.pool.submit(self._line_steps, i, nsteps, what == "both") for i in range(self.NL)]   # line calls contain stream syncs: one host thread each
            for f in futs:
                f.result()
        for c in self.ctx_os[1:] + self.ctx_ls:
            self.ctx_o.wait(c)                  # the first ORB stream's stop event waits for every other stream
        return self.ctx_o.timer_stop()

    # ---- end to end ----
    def prepare_e2e(self):
        torch, B, capk, capl = self.torch, self.B, self.capk, self.capl
        self.h_img = [torch.from_numpy(f).pin_memory() for f in self.frames]
        self.h_kps = torch.empty((B, capk, 28), dtype=torch.uint8).pin_memory(); self.h_desc = torch.empty((B, capk, 32), dtype=torch.uint8).pin_memory()
        self.h_kl = torch.empty((B, capl, 68), dtype=torch.uint8).pin_memory(); self.h_mid = torch.empty((B, capl, 28), dtype=torch.uint8).pin_memory()
        self.h_ld = torch.empty((B, capl, 32), dtype=torch.uint8).pin_memory()
        self.n_k = np.zeros(B, np.int32); self.n_l = np.zeros(B, np.int32)

    def e2e_orb(self, i, k, dimg):
        W, H = self.W, self.H
        for j in range(i * self.BO, (i + 1) * self.BO, self.OSUB):
            s = slice(j, j + self.OSUB)
            self.ctx_os[i].check(self.lib.plf_orb_extract_batch_from_device(self.orbs[i].h, dimg[s].data_ptr(), self.OSUB, W, H, W, W * H,
                                                                            self.h_kps[s].data_ptr(), self.h_desc[s].data_ptr(), self.capk, self.n_k[s].ctypes.data))

    def e2e_line(self, i, k, dimg):
        W, H = self.W, self.H
        for j in range(i * self.BL, (i + 1) * self.BL, self.SUB):
            s = slice(j, j + self.SUB)
            self.ctx_ls[i].check(self.lib.plf_line_extract_batch_from_device(self.les[i].h, dimg[s].data_ptr(), self.SUB, W, H, W, W * H,
                                                                             self.h_kl[s].data_ptr(), self.h_mid[s].data_ptr(), self.h_ld[s].data_ptr(),
                                                                             self.capl, self.n_l[s].ctypes.data))

    def run_e2e(self, nsteps):
        """The same from HOST buffers through the C ABI: every step's images go from pinned host memory to the device ONCE
        (plf_upload on an uploader context, one step ahead, two staging buffers); the ORB and line extractor instances -- one
        host thread each, the reference's ORB thread and line thread (Frame.cc:301-304) -- wait for the upload on their own
        streams (plf_ctx_wait) and return their results to host memory (plf_*_extract_batch_from_device: H2D once + D2H of
        every result inside the timed region)."""
        torch = self.torch
        if not hasattr(self, "d_stage"):
            self.d_stage = [torch.empty((self.B, self.H, self.W), dtype=torch.uint8, device="cuda") for _ in range(2)]
            self.ctx_up = self.S.Context(self.dev)
        torch.cuda.synchronize()
        nbytes = self.B * self.W * self.H
        ncons = self.NO + self.NL
        free = [threading.Semaphore(1), threading.Semaphore(1)]
        uploaded = [threading.Event() for _ in range(nsteps)]
        left = [ncons] * nsteps
        lock = threading.Lock()
        t0 = time.perf_counter()

        abort = threading.Event()

        def guarded(fn):
            # a failing worker must not leave the others waiting for it for ever (they would hold the rank at the next barrier
            # until the NCCL watchdog fires): flag the abort, wake everybody, let the exception travel through the future
            def run(*a):
                try:
                    fn(*a)
                except BaseException:
                    abort.set()
                    for sem in free:
                        sem.release()
                    for ev in uploaded:
                        ev.set()
                    raise
            return run

        def check_abort():
            if abort.is_set():
                raise RuntimeError("another e2e worker failed")

        def uploader():
            for k in range(nsteps):
                free[k % 2].acquire()
                check_abort()
                self.ctx_up.check(self.lib.plf_upload(self.ctx_up.h, self.d_stage[k % 2].data_ptr(), self.h_img[k % self.nsets].data_ptr(), nbytes))
                uploaded[k].set()

        def consumer(kind, i):
            ctx = self.ctx_os[i] if kind == "orb" else self.ctx_ls[i]
            for k in range(nsteps):
                uploaded[k].wait()
                check_abort()
                ctx.wait(self.ctx_up)                      # this stream waits for the upload (device-side dependency, no host sync)
                (self.e2e_orb if kind == "orb" else self.e2e_line)(i, k, self.d_stage[k % 2])
                with lock:
                    left[k] -= 1
                    last = left[k] == 0
                if last:
                    free[k % 2].release()

        futs = [self.pool.submit(guarded(uploader))] + [self.pool.submit(guarded(consumer), "orb", i) for i in range(self.NO)] + \
               [self.pool.submit(guarded(consumer), "line", i) for i in range(self.NL)]
        errs = []
        for f in futs:
            try:
                f.result()
            except BaseException as e:      # collect them all: the first one raised is the cause, the rest are "another worker failed"
                errs.append(e)
        if errs:
            raise [e for e in errs if "another e2e worker failed" not in str(e)][0] if any("another e2e worker failed" not in str(e) for e in errs) else errs[0]
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    def e2e_bytes(self):
        B = self.B
        return B * self.W * self.H, B * (self.capk * 60 + 4 + self.capl * (68 + 28 + 32) + 4)     # every frame is uploaded once

    def check_outputs(self):
        nk = self.d_nk.cpu().numpy(); nl = self.d_nl.cpu().numpy()
        assert (nk > 0).all() and (nl >= 0).all(), "extraction reported an overflow (n_out < 0)"
        return nk, nl

    def checksum(self):
        """Order-sensitive checksum of the step's results on the device: keypoint / keyline records and both descriptor sets
        of every frame (valid rows only)."""
        torch = self.torch
        nk, nl = self.d_nk.long(), self.d_nl.long()
        mk = (torch.arange(self.capk, device="cuda")[None, :] < nk[:, None])
        ml = (torch.arange(self.capl, device="cuda")[None, :] < nl[:, None])
        tot = 0
        for t, m in ((self.d_kps, mk), (self.d_desc, mk), (self.d_kl, ml), (self.d_ld, ml)):
            v = t.to(torch.int64) * m[:, :, None]
            wgt = (torch.arange(v.shape[1] * v.shape[2], device="cuda", dtype=torch.int64) % 65521 + 1).view(1, v.shape[1], v.shape[2])
            tot += int((v * wgt).sum().item())
        return tot & 0xFFFFFFFFFFFF


def algorithmic_bytes(cfg):
    """Algorithmic bytes per frame per kernel: px0 = input pixels, spx = scaled LSD pixels of both
    octaves, sumpx = ORB pyramid pixels; `dens` = fraction of LSD pixels with a defined gradient on this workload."""
    W, H, line, orb = cfg["W"], cfg["H"], cfg["line"], cfg["orb"]
    sw, sh = int(round(W * line["scale"])), int(round(H * line["scale"]))
    spx = sw * sh + (int(round((W // 2) * line["scale"])) * int(round((H // 2) * line["scale"])))
    lv = [(int(np.rint(np.float32(W) / np.float32(orb["scaleFactor"]) ** l)), int(np.rint(np.float32(H) / np.float32(orb["scaleFactor"]) ** l))) for l in range(orb["nlevels"])]
    sumpx = sum(a * b for a, b in lv)
    px0 = W * H
    p01 = px0 + px0 // 4
    dens = 0.07
    alg = {
        "k_lsd_grow_warp": 6 * spx, "k_lsd_grow": 6 * spx,
        "k_lsd_grad": int((1 + 4 + 1 / 8 + dens * 8) * spx), "k_ccl_merge": int((1 / 8 + dens * 8) * spx),
        "k_lsd_keys": int((1 / 8 + dens * 16) * spx), "k_lsd_cid": int(dens * (8 + 4 + 4 + 8) * spx),
        "k_lsd_rect": int(dens * 2 * (4 + 4) * spx),
        "k_fast_cells": sumpx, "k_fast_cells_tma": sumpx, "k_blur7": 2 * sumpx, "k_resize_linear": 2 * sumpx - 2 * px0 + (px0 - lv[-1][0] * lv[-1][1]),
        "k_gauss_strip<3>": 2 * p01, "k_gauss_strip<2>": 2 * px0, "k_resize_exact": p01 + spx,
        "k_pyrdown": 2 * (px0 + px0 // 4), "k_sobel3": 5 * p01,
        # own radix sort: 4 passes over the 8-byte keys of the defined pixels (histogram reads them once, scatter reads and writes)
        "k_rs_hist": int(4 * 8 * dens * spx), "k_rs_scatter": int(4 * 16 * dens * spx), "k_scan_apply": int(spx / 32 * 8), "k_scan_tile_sums": int(spx / 32 * 4),
        "k_describe": 2 * 1024 * orb["nfeatures"], "k_describe_tma": (48 * 31 + 64 * 39 + 60) * orb["nfeatures"], "k_lbd": 63 * 4 * 60 * line["nfeatures"],
    }
    # B_orb + the per-octave line figures (LSD-pre, gradient, sort, grow, LBD-prep, LBD gathers)
    b_orb = px0 + sumpx + (sumpx - lv[-1][0] * lv[-1][1]) + sumpx + 2 * sumpx + 60 * orb["nfeatures"]
    b_line = 0
    for p in (px0, px0 // 4):
        s2p = line["scale"] ** 2 * p
        b_line += (2 * p + p + s2p) + 9 * s2p + 2 * 8 * s2p * dens + 6 * s2p + (2 * p + (p + p / 4) + 5 * p)
    b_line += 63 * 4 * 60 * line["nfeatures"]
    return alg, int(b_orb), int(b_line), spx, sumpx


def concurrent_profile(ex, nst):
    """Per-kernel time while ALL streams run together, as in the timed region: CUDA-event brackets around every launch
    (plf_profile_enable), each kernel's busy time = the union of its launch intervals over `nst` steps."""
    for c in ex.contexts():
        c.profile_enable(True)
    t = ex.run_device(nst) / nst
    per = {}
    for c in ex.contexts():
        for name, t0, t1 in c.profile_timeline(ex.ctx_o):
            per.setdefault(name, []).append((t0, t1))
        c.profile_enable(False)

    def union(iv):
        tot, c0, c1 = 0.0, None, None
        for a, b in sorted(iv):
            if c1 is None or a > c1:
                if c1 is not None:
                    tot += c1 - c0
                c0, c1 = a, b
            else:
                c1 = max(c1, b)
        return tot + ((c1 - c0) if c1 is not None else 0.0)

    out = {k: {"busy_ms": union(v) / nst, "sum_ms": sum(b - a for a, b in v) / nst, "launches": len(v) // nst} for k, v in per.items()}
    allk = union([iv for v in per.values() for iv in v]) / nst
    return t, out, allk


def measure_extraction(S, torch, dist, name, cfg, args, rank, world, dev, barrier, maxr, headline):
    """Device-resident value, e2e, ORB-only, roofline for one extraction config.  Returns the pieces of the JSON line."""
    B = args.frames or cfg["frames"]
    if not headline:
        B = min(B, 256)
    steps, warm = (args.steps, max(args.warmup, 3)) if headline else (3, 3)
    ex = Extraction(S, torch, cfg, B, rank, world, dev, args)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    ex.run_device(warm)
    flush.zero_()
    barrier()
    sampler = ClockSampler(dev) if (rank == 0 and headline) else None
    l0 = sum(c.launch_count() for c in ex.contexts())
    ms_dev = ex.run_device(steps)
    barrier()
    launches = sum(c.launch_count() for c in ex.contexts()) - l0
    clocks = sampler.stop() if sampler else None
    nk, nl = ex.check_outputs()
    csum = ex.checksum()
    # ORB only
    ex.run_device(2, "orb"); flush.zero_(); barrier()
    ms_orb = ex.run_device(steps, "orb")
    barrier()
    ms_line = None
    if headline:
        ex.run_device(1, "line"); flush.zero_(); barrier()
        ms_line = ex.run_device(steps, "line")
        barrier()
    prof = None
    if rank == 0 and headline:
        prof = concurrent_profile(ex, 3)
    barrier()
    # end to end
    ex.prepare_e2e()
    ex.run_e2e(warm if headline else 1)
    flush.zero_()
    barrier()
    ms_e2e = ex.run_e2e(steps)
    barrier()
    h2d, d2h = ex.e2e_bytes()
    ms_dev, ms_e2e, ms_orb = maxr(ms_dev), maxr(ms_e2e), maxr(ms_orb)
    if ms_line is not None:
        ms_line = maxr(ms_line)
    res = dict(B=B, steps=steps, warm=warm, ms_dev=ms_dev, ms_e2e=ms_e2e, ms_orb=ms_orb, ms_line=ms_line, launches=int(launches), clocks=clocks,
               nk=float(nk.mean()), nl=float(nl.mean()), checksum=csum, h2d=h2d, d2h=d2h, prof=prof, NO=ex.NO, NL=ex.NL, SUB=ex.SUB, OSUB=ex.OSUB,
               sample=ex.frames[0][:min(B, 512)].copy() if rank == 0 else None)
    ex.close()
    del flush
    torch.cuda.empty_cache()
    return res


def latency_c1(S, cfg, dev, nrep=60):
    """BASELINE config 1 / the drop-in use: one frame through plf_orb_extract || plf_line_extract from host buffers, the
    reference's two threads (src/Frame.cc:301-304).  Returns median / p90 milliseconds per frame."""
    W, H, orb, line = cfg["W"], cfg["H"], cfg["orb"], cfg["line"]
    imgs = [synth_image(W, H, s) for s in range(4)]
    co, cl = S.Context(dev), S.Context(dev, priority=1)
    eo = S.ORBextractor(orb["nfeatures"], orb["scaleFactor"], orb["nlevels"], orb["iniThFAST"], orb["minThFAST"], ctx=co)
    el = S.Lineextractor(line["nfeatures"], line["nlevels"], line["refine"], line["scale"], line["sigma_scale"], line["quant"], line["ang_th"],
                         line["log_eps"], line["density_th"], line["n_bins"], line["min_line_length"], ctx=cl)
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(2)
    ts, to, tl = [], [], []

    def timed(fn, img):
        t0 = time.perf_counter(); fn(img); return (time.perf_counter() - t0) * 1e3

    for r in range(nrep + 5):
        img = imgs[r % len(imgs)]
        t0 = time.perf_counter()
        fo = pool.submit(timed, eo, img); fl = pool.submit(timed, el.ComputeLsdWithLbd, img)
        a, b = fo.result(), fl.result()
        dt = (time.perf_counter() - t0) * 1e3
        if r >= 5:
            ts.append(dt); to.append(a); tl.append(b)
    pool.shutdown(); eo.close(); el.close(); co.close(); cl.close()
    return {"frame_ms_median": float(np.median(ts)), "frame_ms_p90": float(np.percentile(ts, 90)), "orb_ms_median": float(np.median(to)),
            "line_ms_median": float(np.median(tl)), "frames_per_s": 1e3 / float(np.median(ts)), "reps": nrep,
            "how": "plf_orb_extract || plf_line_extract on two host threads, host buffers, wall clock per frame"}


def stereo_c3(S, torch, cfg, dev, npairs=128, reps=5):
    """BASELINE config 3's matching half on device-resident extraction results: Frame::ComputeStereoMatches for `npairs` pairs in
    one call (plf_stereo_match_batch_device) and the L<->R line matchNNR of every pair (knn2 + ratio test, nnr 0.75)."""
    W, H, orb, line = cfg["W"], cfg["H"], cfg["orb"], cfg["line"]
    frames = make_frames(cfg, list(range(2 * npairs)))
    ctx = S.Context(dev)
    lib = ctx.lib
    ex = S.ORBextractor(orb["nfeatures"], orb["scaleFactor"], orb["nlevels"], orb["iniThFAST"], orb["minThFAST"], ctx=ctx)
    le = S.Lineextractor(line["nfeatures"], line["nlevels"], line["refine"], line["scale"], line["sigma_scale"], line["quant"], line["ang_th"],
                         line["log_eps"], line["density_th"], line["n_bins"], line["min_line_length"], ctx=ctx)
    capk, capl, nfr = ex.max_keypoints, le.max_keylines, 2 * npairs
    d_img = torch.from_numpy(frames).cuda()
    dk = torch.empty((nfr, capk, 28), dtype=torch.uint8, device="cuda"); dd = torch.empty((nfr, capk, 32), dtype=torch.uint8, device="cuda")
    dn = torch.empty(nfr, dtype=torch.int32, device="cuda")
    du = torch.empty((npairs, capk), dtype=torch.float32, device="cuda"); dz = torch.empty((npairs, capk), dtype=torch.float32, device="cuda")
    kl = torch.empty((nfr, capl, 68), dtype=torch.uint8, device="cuda"); mid = torch.empty((nfr, capl, 28), dtype=torch.uint8, device="cuda")
    ld = torch.empty((nfr, capl, 32), dtype=torch.uint8, device="cuda"); nl = torch.empty(nfr, dtype=torch.int32, device="cuda")
    ctx.check(lib.plf_orb_extract_batch_device(ex.h, d_img.data_ptr(), nfr, W, H, W, W * H, dk.data_ptr(), dd.data_ptr(), capk, dn.data_ptr()))
    ctx.check(lib.plf_line_extract_batch_device(le.h, d_img.data_ptr(), nfr, W, H, W, W * H, kl.data_ptr(), mid.data_ptr(), ld.data_ptr(), capl, nl.data_ptr()))
    ctx.synchronize()
    nlh = nl.cpu().numpy()
    mb, mbf = 0.54, 0.54 * 718.856
    idx = torch.empty((capl, 2), dtype=torch.int32, device="cuda"); dst = torch.empty((capl, 2), dtype=torch.int32, device="cuda")
    m12 = torch.empty(capl, dtype=torch.int32, device="cuda"); nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    ts, tl = [], []
    for r in range(reps + 1):
        ctx.timer_start()
        ctx.check(lib.plf_stereo_match_batch_device(ex.h, ex.h, npairs, 0, 2, 1, 2, dk.data_ptr(), dd.data_ptr(), dn.data_ptr(), dk.data_ptr(),
                                                    dd.data_ptr(), dn.data_ptr(), capk, mb, mbf, du.data_ptr(), dz.data_ptr()))
        a = ctx.timer_stop()
        ctx.timer_start()
        for p_ in range(npairs):
            nq, nt = int(nlh[2 * p_]), int(nlh[2 * p_ + 1])
            if nq == 0 or nt < 2:
                continue
            ctx.check(lib.plf_hamming_knn2_device(ctx.h, ld[2 * p_].data_ptr(), nq, ld[2 * p_ + 1].data_ptr(), nt, 0, idx.data_ptr(), dst.data_ptr()))
            ctx.check(lib.plf_nnr_from_knn2_device(ctx.h, idx.data_ptr(), dst.data_ptr(), nq, 0.75, m12.data_ptr(), nm.data_ptr()))
        b = ctx.timer_stop()
        if r:
            ts.append(a); tl.append(b)
    matched = float((du[:, :] >= 0).sum().item()) / npairs
    ex.close(); le.close(); ctx.close()
    return {"pairs": npairs, "stereo_points_ms": float(np.median(ts)), "stereo_pairs_per_s": npairs / (float(np.median(ts)) / 1e3),
            "line_nnr_ms": float(np.median(tl)), "line_nnr_pairs_per_s": npairs / (float(np.median(tl)) / 1e3), "mean_stereo_matches": matched,
            "what": "plf_stereo_match_batch_device over all pairs in one call; per pair plf_hamming_knn2_device + plf_nnr_from_knn2_device on the line descriptors"}


def measure_matching(S, torch, dist, args, rank, world, dev, barrier, maxr, sizes, cpu=True):
    """BASELINE config 5: 1e4 queries x T train rows, train-sharded across ranks with the NCCL merge; every size is also
    checked against the unsharded result through a checksum of the merged table (N > 1)."""
    from spl_slam_b200 import sharded
    NQ = 10000
    ctx_m = S.Context(dev)
    comm = sharded.Comm(ctx_m)          # C-ABI communicator (plf_comm_*): NCCL all-gather + merge queued by libplf.so
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    gq = torch.Generator(device="cuda"); gq.manual_seed(1234)
    dq = torch.randint(0, 256, (NQ, 32), dtype=torch.uint8, device="cuda", generator=gq)
    out = []
    popc_peak = ctx_m.popc_peak()
    CH = 100000

    def train_rows(lo, hi):
        """Rows [lo, hi) of the global train set: fixed chunks of 1e5 rows seeded by the chunk number, so every world size
        (and the unsharded check) sees the SAME set."""
        parts = []
        for c in range(lo // CH, (hi + CH - 1) // CH):
            g = torch.Generator(device="cuda"); g.manual_seed(4321 + c)
            blk = torch.randint(0, 256, (CH, 32), dtype=torch.uint8, device="cuda", generator=g)
            parts.append(blk[max(lo, c * CH) - c * CH:min(hi, (c + 1) * CH) - c * CH])
        return torch.cat(parts) if parts else torch.empty((0, 32), dtype=torch.uint8, device="cuda")

    for NT in sizes:
        tb, te = sharded.train_shard(NT, rank, world)
        dt = train_rows(tb, te)
        for _ in range(2):
            idx, dst, m12, nm = comm.match_nnr(dq, dt, tb, 0.75)
        ctx_m.synchronize()
        barrier()
        reps = []
        for _ in range(5 if NT >= 10 ** 7 else 7):
            flush.zero_(); torch.cuda.synchronize()
            ctx_m.timer_start()
            idx, dst, m12, nm = comm.match_nnr(dq, dt, tb, 0.75)       # local top-2 + ncclAllGather + merge + ratio test, one stream
            reps.append(ctx_m.timer_stop())
        barrier()
        ms = maxr(float(np.median(reps)))
        verified = None
        if world > 1 and NT <= 10 ** 6:
            # the NCCL-merged table must equal the unsharded one: every rank searches the whole train set itself and compares
            full = train_rows(0, NT)
            ui = torch.empty((NQ, 2), dtype=torch.int32, device="cuda"); ud = torch.empty((NQ, 2), dtype=torch.int32, device="cuda")
            ctx_m.check(ctx_m.lib.plf_hamming_knn2_device(ctx_m.h, dq.data_ptr(), NQ, full.data_ptr(), NT, 0, ui.data_ptr(), ud.data_ptr()))
            ctx_m.synchronize()
            assert torch.equal(ui, idx) and torch.equal(ud, dst), "sharded top-2 differs from the unsharded table at %d rows" % NT
            verified = "equal to the unsharded table on every rank"
            del full
        # checksum of the merged top-2 + matches: equal for every world size (same global train set), reported per N
        wq = (torch.arange(NQ, device="cuda", dtype=torch.int64) % 65521 + 1)
        csum = int(((idx.to(torch.int64) * 3 + dst.to(torch.int64)).sum(1) * wq).sum().item() + (m12.to(torch.int64) * wq).sum().item()) & 0xFFFFFFFFFFFF
        ent = {"train_rows": NT, "ms": ms, "queries_per_s": NQ / (ms / 1e3), "pairs_per_s": NQ * NT / (ms / 1e3), "matches": int(nm.item()), "checksum": csum,
               "popc_frac": 8 * NQ * NT / (ms / 1e3) / world / popc_peak}
        if verified:
            ent["verified"] = verified
        out.append(ent)
        del dt
    comm.close()
    ctx_m.close()
    return out, popc_peak


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (0 = the configuration's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline configuration only")
    ap.add_argument("--line-contexts", type=int, default=LINE_CONTEXTS)
    ap.add_argument("--orb-contexts", type=int, default=ORB_CONTEXTS)
    ap.add_argument("--line-sub", type=int, default=0, help="frames per line call (0 = the configuration's)")
    ap.add_argument("--orb-sub", type=int, default=0, help="frames per ORB call (0 = min(share, 256))")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: everything else that libraries print to file descriptor 1 (NCCL's version
    # banner, for one) goes to stderr until the result line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import spl_slam_b200 as S
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a rank that dies must take the job down quickly, not after NCCL's default 10-minute watchdog
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=240))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev = local_rank
    name = args.config
    t_start = time.time()

    def emit(line):
        if rank == 0:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)

    if name == "c5":
        sweep, popc_peak = measure_matching(S, torch, dist, args, rank, world, dev, barrier, maxr, [10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7])
        e6 = [e for e in sweep if e["train_rows"] == 10 ** 6][0]
        line = {"metric": "Hamming matches/s", "value": e6["queries_per_s"], "unit": "queries/s at 1e6 train rows", "n_gpus": world, "steps": 7,
                "warmup": 2, "ms_per_step": e6["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": c5_config(world, 10 ** 6), "sweep": sweep,
                "roofline": {"bound": "popc", "achieved": 8 * e6["pairs_per_s"] / world, "peak": popc_peak, "unit": "popc32/s per GPU",
                             "frac": e6["popc_frac"], "peak_source": "plf_popc_peak micro-benchmark on this GPU"}}
        emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    cfg = CONFIGS[name]
    if name == "c1":
        lat = latency_c1(S, cfg, dev, nrep=max(args.steps, 20) * 5)
        line = {"metric": "frames/s ORB+LSD/LBD extraction", "value": lat["frames_per_s"], "unit": "frames/s", "n_gpus": 1, "steps": lat["reps"],
                "warmup": 5, "ms_per_step": lat["frame_ms_median"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config_dict(name, cfg, 1, 1), "latency": lat}
        emit(line)
        return

    r = measure_extraction(S, torch, dist, name, cfg, args, rank, world, dev, barrier, maxr, True)
    B = r["B"]
    total_frames = B * world * r["steps"]
    value = total_frames / (r["ms_dev"] / 1e3)
    e2e_value = total_frames / (r["ms_e2e"] / 1e3)

    others = {}
    if not args.no_extras and name == "c4":
        # short measurements of the other BASELINE configurations (extra keys; the headline stays c4)
        for oname in ("c2", "c3"):
            o = measure_extraction(S, torch, dist, oname, CONFIGS[oname], args, rank, world, dev, barrier, maxr, False)
            tf = o["B"] * world * o["steps"]
            others[oname] = {"workload": CONFIGS[oname]["workload"], "frames_per_gpu_per_step": o["B"], "steps": o["steps"],
                             "frames_per_s": tf / (o["ms_dev"] / 1e3), "ms_per_step": o["ms_dev"] / o["steps"],
                             "e2e_frames_per_s": tf / (o["ms_e2e"] / 1e3), "orb_only_frames_per_s": tf / (o["ms_orb"] / 1e3),
                             "mean_keypoints": o["nk"], "mean_lines": o["nl"], "checksum": o["checksum"]}
        if rank == 0:
            others["c1"] = dict(latency_c1(S, CONFIGS["c1"], dev), workload=CONFIGS["c1"]["workload"])
            c3 = CONFIGS["c3"]
            others["c3_single_image"] = dict(latency_c1(S, c3, dev, nrep=30), workload="one 1241x376 image through the drop-in calls (per image)")
            others["c3"]["stereo_matching"] = stereo_c3(S, torch, c3, dev)
        barrier()
        sweep, popc_peak = measure_matching(S, torch, dist, args, rank, world, dev, barrier, maxr, [10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7])
        others["c5"] = {"workload": c5_config(world, 10 ** 6)["workload"].replace("1e+06", "1e4..1e7"), "sweep": sweep, "popc_peak": popc_peak,
                        "popc_peak_source": "plf_popc_peak micro-benchmark on this GPU"}

    if rank == 0:
        peak, peak_src = peaks()
        alg, b_orb, b_line, spx, sumpx = algorithmic_bytes(cfg)
        ms_step = r["ms_dev"] / r["steps"]
        t_conc, prof, busy_any = r["prof"] if r["prof"] else (None, {}, None)
        roof = None
        if prof:
            # the dominant kernel = the one whose launches cover the largest part of the step while everything runs together
            name_k, pk = max(prof.items(), key=lambda kv: kv[1]["busy_ms"])
            nl_ = max(pk["launches"], 1)
            ach = alg.get(name_k, 0) * B / nl_ / (pk["sum_ms"] / nl_ / 1e3) / 1e9 if pk["sum_ms"] > 0 else 0.0
            traffic, traffic_src = None, None
            tp = os.path.join(ROOT, "profiles", "r2_dominant_kernel_ncu.json")
            if os.path.exists(tp):
                try:
                    tj = json.load(open(tp))
                    if tj.get("kernel") == name_k:      # per-frame DRAM bytes of the capture x the frames of one launch here
                        traffic = tj.get("dram_bytes_per_frame", 0) * B / nl_ * (2 if name_k.startswith("k_lsd") else 1)
                        traffic_src = tj.get("note")
                except Exception:
                    pass
            roof = {"bound": "hbm", "kernel": name_k, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": traffic_src, "peak_source": peak_src, "launches_per_step": pk["launches"],
                    "avg_launch_ms": pk["sum_ms"] / nl_, "busy_ms_per_step": pk["busy_ms"], "share_of_step": pk["busy_ms"] / t_conc,
                    "algorithmic_bytes_per_frame": alg.get(name_k, 0),
                    "how": "CUDA-event brackets around every launch while all streams run together (the timed configuration, %.1f ms/step "
                           "with the brackets on); busy = union of the kernel's launch intervals, so no kernel exceeds the step" % t_conc,
                    "whole_step": {"algorithmic_bytes_per_frame": b_orb + b_line, "orb_bytes_per_frame": b_orb, "line_bytes_per_frame": b_line,
                                   "achieved_gbs": (b_orb + b_line) * B / (ms_step / 1e3) / 1e9, "frac": (b_orb + b_line) * B / (ms_step / 1e3) / 1e9 / peak,
                                   "orb_only_gbs": b_orb * B * r["steps"] / (r["ms_orb"] / 1e3) / 1e9,
                                   "orb_only_frac": b_orb * B * r["steps"] / (r["ms_orb"] / 1e3) / 1e9 / peak,
                                   "note": "SURVEY.md 8d byte counts (B_orb + per-octave line figures) x frames / device step time"},
                    "busy_any_kernel_ms_per_step": busy_any,
                    "kernels": [{"kernel": k, "busy_ms_per_step": round(v["busy_ms"], 4), "sum_ms_per_step": round(v["sum_ms"], 4),
                                 "launches_per_step": v["launches"], "algorithmic_bytes_per_frame": alg.get(k),
                                 "achieved_gbs": round(alg[k] * B / (v["sum_ms"] / 1e3) / 1e9, 1) if alg.get(k) and v["sum_ms"] > 0 else None,
                                 "frac": round(alg[k] * B / (v["sum_ms"] / 1e3) / 1e9 / peak, 4) if alg.get(k) and v["sum_ms"] > 0 else None}
                                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["busy_ms"])]}
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            per_frame_s = 1.1e-6 * cfg["W"] * cfg["H"] * 0.25
            ns = max(cores, min(len(r["sample"]), int(15.0 * cores / per_frame_s)))
            sample = r["sample"][:ns]
            fps, dt, kind = cpu_reference_throughput(cfg, sample, cores)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                   "sample": "%d of the step's frames, one frame per thread on %d threads, %.1f s (%s)" %
                             (len(sample), cores, dt, "oracle/_ref: the reference's own sources compiled unmodified" if kind == "reference" else "C oracle")}
            if others.get("c5"):
                rng = np.random.default_rng(1234)
                q = rng.integers(0, 256, (10000, 32), dtype=np.uint8); t = rng.integers(0, 256, (100000, 32), dtype=np.uint8)
                t1 = cpu_knn2(q[:2000], t, 1); tn = cpu_knn2(q, t, cores)
                others["c5"]["cpu_baseline"] = {"kind": "reference", "what": "cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) (src/Linematcher.cc:526-527)",
                                                "pairs_per_s_1_thread": 2000 * 100000 / t1, "pairs_per_s_all_threads": 10000 * 100000 / tn, "cores": cores,
                                                "sample": "1e4 (2e3 for one thread) queries x 1e5 train rows"}
        line = {"metric": "frames/s ORB+LSD/LBD extraction", "value": value, "unit": "frames/s", "n_gpus": world, "steps": r["steps"],
                "warmup": r["warm"], "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_dict(name, cfg, B, world),
                "run": {"orb_contexts": r["NO"], "line_contexts": r["NL"], "frames_per_line_call": r["SUB"], "frames_per_orb_call": r["OSUB"],
                        "distinct_frame_sets": NSETS,
                        "l2": "inputs larger than L2: %d distinct frames (%.1f GB) per GPU, step k works on set k %% %d; 256 MiB flush before the timed "
                              "region; the K steps run back to back, no barrier between steps" % (NSETS * B, NSETS * B * cfg["W"] * cfg["H"] / 1e9, NSETS)},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "ms_per_step": r["ms_e2e"] / r["steps"]},
                "orb_only": {"value": total_frames / (r["ms_orb"] / 1e3), "unit": "frames/s", "ms_per_step": r["ms_orb"] / r["steps"]},
                "lines_only": {"value": total_frames / (r["ms_line"] / 1e3), "unit": "frames/s", "ms_per_step": r["ms_line"] / r["steps"]} if r["ms_line"] else None,
                "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": roof, "cpu_baseline": cpu,
                "outputs": {"mean_keypoints": r["nk"], "mean_lines": r["nl"], "checksum_rank0": r["checksum"]},
                "other_configs": others or None, "bench_wall_s": round(time.time() - t_start, 1)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
