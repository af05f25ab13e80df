#!/usr/bin/env python
"""bench.py -- images/sec of the segment + CC hot path (BASELINE.json metric) on N B200s.

A step = one pass of the hot path (net forward, logit threshold, connected components, boxes) over
one batch of synthetic 1024x1024 grayscale images per GPU (BASELINE configs[1]: batch 64).
  value     inputs resident in HBM, device-timed with CUDA events on the launching stream,
            max over ranks, whole-job images/sec
  e2e       same metric through the reference-facing C-ABI call (ubd_segment) with pinned HOST
            buffers: H2D of the images and D2H of mask + components inside the timed region
  roofline  the dominant kernel (dilated 3x3 conv): algorithmic FLOP / launch time vs the measured bf16 tensor
            peak (tf32 = 1/2); stem, CC and whole-step fractions beside it
  cpu_baseline  the CPU oracle (torch-CPU restatement of the reference + its cv2 calls) on a bounded
            sample, rank 0, N=1 only
`--impl reference` times that CPU restatement as the reference arm (TF/Keras cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec at 1024x1024 (net+CC postproc)"
ALGO_MAC_PER_INPUT_PX = 2201.25          # whole forward, C = 0
DIL_MAC_PER_MAP_PX = 5184                # one dilated 3x3 24->24 layer
ALGO_BYTES_PER_IMAGE_1024 = 4784128      # SURVEY.md 8d (f32 in + logits + mask + labels)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons DURING the timed regions (B200_PROFILING.md).  A thread polls NVML
    (nvidia_ml_py) every ~2 ms from before the warm-up until after the end-to-end loop; `mark()` brackets
    the timed regions so that `sm_mhz` is the median over samples taken inside them.  Falls back to one
    background `nvidia-smi -lms 100` process when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []
        self.samples, self.windows, self._t0 = [], [], None       # (t, sm_mhz, reasons_bitmask, power_w)
        self.nvml, self.thread, self.stop = None, None, False
        self.sm_max = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.replace(",", "").isdigit() else self.index
            self.handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
            self.nvml = nv
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            except Exception:
                self.proc = None
        return self

    def mark(self, begin):
        """Bracket a timed region (wall clock; the regions are synchronised on both sides)."""
        if begin:
            self._t0 = time.perf_counter()
        elif self._t0 is not None:
            self.windows.append((self._t0, time.perf_counter()))
            self._t0 = None

    def __exit__(self, *a):
        if self.thread is not None:
            self.stop = True
            self.thread.join(timeout=2)
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        self.rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.count(",") >= 8]

    def _summary_nvml(self):
        nv = self.nvml
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        pmax = max((s[3] for s in self.samples), default=0.0)
        loaded = inside if inside else [s for s in self.samples if s[3] >= 0.5 * pmax]
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, v in bits.items() if any(s[2] & v for s in loaded))
        return {"sm_mhz": statistics.median(s[1] for s in loaded) if loaded else None, "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(self.samples), "samples_in_timed_regions": len(inside),
                "power_w_max": pmax, "source": "nvml"}

    def summary(self):
        if self.nvml is not None and self.samples:
            return self._summary_nvml()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None
        power = [num(r[3]) for r in self.rows]
        sm_all = [num(r[1]) for r in self.rows]
        # "under load": samples drawing more than half of the highest power seen
        pmax = max([p for p in power if p is not None], default=0.0)
        sm = [s for s, p in zip(sm_all, power) if s is not None and (p is None or p >= 0.5 * pmax)]
        mx = [num(r[2]) for r in self.rows if num(r[2]) is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nme, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "samples_under_load": len(sm),
                "power_w_max": pmax, "source": "nvidia-smi"}


def make_images(batch, h, w, seed=1):
    """`batch` uint8 images of h x w: 8 distinct synthetic images, tiled (synthetic data)."""
    from ubdvss_b200 import synth
    base = synth.synth_images(min(batch, 8), h, w, seed=seed)
    reps = -(-batch // base.shape[0])
    return np.ascontiguousarray(np.concatenate([base] * reps, 0)[:batch])


def cpu_reference_step(weights, images_u8, thr, n_classes):
    """One pass of the CPU restatement of the reference over `images_u8` (oracle, test infra)."""
    from oracle import net as onet, postproc as pp
    x = onet.preprocess(images_u8.astype(np.float64), "mobilenet_like").astype(np.float32)
    logits = onet.forward_torch(weights, x)
    det = pp.threshold_mask(logits[..., :1], thr)
    return [pp.postprocess_cv2(det[i], logits[i, ..., 1:] if n_classes else None, 4, 5) for i in range(det.shape[0])]


def time_cpu(weights, images_u8, thr, n_classes, steps, warmup):
    import torch
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    for _ in range(warmup):
        cpu_reference_step(weights, images_u8, thr, n_classes)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(weights, images_u8, thr, n_classes)
    dt = time.perf_counter() - t0
    return images_u8.shape[0] * steps / dt, dt / steps, cores


def measure_traffic(args, precision):
    """(DRAM bytes per launch of the dilated-conv kernel, DRAM bytes of the whole step, source) from an ncu pass over one
    step of this very configuration, or (None, None, reason)."""
    import csv
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, None, "ncu not found"
    log = tempfile.NamedTemporaryFile(suffix=".csv", delete=False).name
    cmd = [ncu, "--profile-from-start", "off", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "--cache-control", "none", "--csv", "--log-file", log, sys.executable, os.path.abspath(__file__), "--traffic-child",
           "--config", args.config, "--precision", precision, "--no-train", "--no-cpu-baseline"]
    if args.batch:
        cmd += ["--batch", str(args.batch)]
    if args.size:
        cmd += ["--size", str(args.size)]
    try:
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600, check=True)
        per = {}
        for r in csv.reader(open(log)):
            if len(r) > 14 and r[0].isdigit() and r[12].startswith("dram__bytes"):
                v = float(r[14].replace(",", ""))
                unit = r[13].lower()
                v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
                k = per.setdefault(int(r[0]), [r[4], 0.0])
                k[1] += v
        if not per:
            return None, None, "ncu produced no kernel records"
        step = sum(v for _, v in per.values())
        import re
        dil = []
        for name, v in per.values():              # the dilated layers: dilconv_col_kernel<BF16, L1SRC = 0, PIPE>
            mt = re.search(r"dilconv_col_kernel<\D*(\d)\D+(\d)", name)
            if mt and mt.group(2) == "0":
                dil.append(v)
        per_launch = sum(dil) / len(dil) if dil else None
        return per_launch, step, f"ncu pass in this run ({len(per)} kernel launches of one step, --cache-control none)"
    except Exception as e:      # noqa: BLE001
        return None, None, f"ncu pass failed: {type(e).__name__}"
    finally:
        try:
            os.unlink(log)
        except OSError:
            pass


def pin_to_gpu_cpus(index):
    """Restrict this rank to the CPUs NVML reports as local to its GPU (pinned staging buffers are then first-touched on
    that NUMA node).  Returns a short description for the JSON line; any failure leaves the affinity alone."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.replace(",", "").isdigit() else index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        cur = os.sched_getaffinity(0)
        want = cpus & cur
        if want and want != cur:
            os.sched_setaffinity(0, want)
        return f"{len(want or cur)} of {len(cur)} CPUs (NVML affinity of GPU {phys})"
    except Exception as e:      # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (NCCL's version banner, torchrun notices) goes to stderr: fd 1 is pointed at fd 2 and
    the ONE JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


# BASELINE.json configs -> (batch per GPU, H, W, precision, classes).  C: 3840x2160 scans are fed as 2176x3840
# (sides rounded up to multiples of 64, segmap_manager.py:153-165); D: 256 images over 8 GPUs = 32 per GPU.
CONFIGS = {
    "B": dict(idx=1, batch=64, h=1024, w=1024, precision="tf32", n_classes=0,
              text="batch-64 synthetic 1024x1024 grayscale inference per GPU, fp32/TF32, threshold + CC boxes"),
    "C": dict(idx=2, batch=8, h=2176, w=3840, precision="tf32", n_classes=0,
              text="batch-sharded 3840x2160 document scans (network input 2176x3840), 8 per GPU, threshold + CC boxes"),
    "D": dict(idx=3, batch=32, h=1024, w=1024, precision="bf16", n_classes=26,
              text="bf16 inference with the 26-type barcode head (detection map + per-pixel class vote), batch 256 over 8 GPUs = 32 per GPU, 1024x1024"),
}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS), help="BASELINE.json configs[1] (default), [2] or [3]")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the config's)")
    ap.add_argument("--size", type=int, default=0, help="square image side (default: the config's shape)")
    ap.add_argument("--precision", default=os.environ.get("UBD_PRECISION", ""), choices=["", "fp32", "tf32", "bf16", "f16"],
                    help="default: the config's (configs[1] is quoted as fp32/TF32: tf32 tensor-core path)")
    ap.add_argument("--cpu-sample", type=int, default=4, help="images per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the config-E training-step measurement")
    ap.add_argument("--sync-api", action="store_true", help="time the synchronous ubd_segment[_dev] calls instead of submit/wait")
    ap.add_argument("--no-extra", action="store_true", help="skip the f16 / bf16 container lines beside a tf32 run")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu pass that measures the DRAM bytes of one step")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from ubdvss_b200 import synth as usynth

    C = CONFIGS[args.config]
    B = args.batch or C["batch"]
    H, W = (args.size, args.size) if args.size else (C["h"], C["w"])
    precision = args.precision or C["precision"]
    n_classes = C["n_classes"]
    # random init of the architecture, scaled layer by layer so that the logit map follows the image content
    weights = usynth.synth_weights(n_classes, seed=1234, calibrated=True)
    thr = 0.0                                                  # pixel_threshold 0.5 (model_runner.py:37-38)
    cfg = {"workload": f"configs[{C['idx']}]: {C['text']}" + ("" if (B, H, W, precision) == (C["batch"], C["h"], C["w"], C["precision"])
                                                              else f" [overridden: batch {B}, {H}x{W}, {precision}]"),
           "batch_per_gpu": B, "image": [H, W, 1], "n_classes": n_classes,
           "input_dtype": "uint8 (mobilenet_like preprocessing folded into L1)", "precision": precision,
           "weights": "random init (Glorot) + layer-sequential unit-variance scaling; pixel_threshold 0.5, min_area 5",
           "parallelism": f"batch-sharded x{world}, no collective",
           "l2": f"inputs larger than L2 ({B * H * W / 2**20:.0f} MiB uint8 per batch, two batches alternate; "
                 f"{B * (H // 4) * (W // 4) * 96 * 2 / 2**20:.0f} MiB of maps per sweep)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        imgs = make_images(args.cpu_sample, H, W)
        v, step_s, cores = time_cpu(weights, imgs, thr, n_classes, max(1, args.steps), max(1, min(args.warmup, 1)))
        cfg["reference_arm_batch"] = (f"{args.cpu_sample} images per step (bounded sample of the {B}-image workload; the per-image "
                                      "rate is what is compared)")
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port",
                                 "sample": f"{args.cpu_sample} images of {H}x{W} per step (torch-CPU "
                                           "restatement of net.py + the reference's cv2 post-processing; TF/Keras absent)"},
                "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    from ubdvss_b200 import _lib
    from ubdvss_b200.engine import Engine
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = pin_to_gpu_cpus(local_rank)         # before any pinned allocation: first touch lands on the GPU's NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    eng = Engine(device=local_rank, precision=precision, n_classes=n_classes)
    eng.set_weights(weights)
    for opt in ("tc_variant", "dense_l2", "stem_chunk", "stem_variant", "chunk", "fused_ccl", "gpu_boxes", "pipeline", "pipe_ring", "cc_stream"):
        if os.environ.get("UBD_" + opt.upper()):
            eng.set_option(opt, int(os.environ["UBD_" + opt.upper()]))
    # two different batches alternate, so that nothing of step k is still in L2 for step k+1
    imgs = [make_images(B, H, W, seed=1 + 2 * rank + j) for j in range(2)]
    pinned = [torch.from_numpy(a).pin_memory() for a in imgs]
    h_imgs = [p.numpy() for p in pinned]
    d_imgs = [p.cuda(non_blocking=False) for p in pinned]
    min_area_x2 = 10
    cap = 256 * B
    DEPTH = max(1, min(3, int(os.environ.get("UBD_DEPTH", "3"))))      # batches in flight through submit / wait
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    if args.traffic_child:
        # profiled by the parent under ncu: two warm-up steps, then ONE step between cudaProfilerStart / Stop
        for k in range(3):
            if k == 2:
                torch.cuda.synchronize()
                torch.cuda.cudart().cudaProfilerStart()
            eng.segment_dev(d_imgs[k & 1].data_ptr(), _lib.UBD_U8, B, H, W, thr, min_area_x2, _lib.PREPROC_MOBILENET, max_comps=cap)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return 0

    def run_dev(steps, eng=eng):
        """`steps` batches through the device-resident entry points; returns the last batch's component counts."""
        counts = None
        if args.sync_api:
            for k in range(steps):
                _, counts = eng.segment_dev(d_imgs[k & 1].data_ptr(), _lib.UBD_U8, B, H, W, thr, min_area_x2, _lib.PREPROC_MOBILENET,
                                            max_comps=cap)
            return counts
        pend = []
        for k in range(steps):
            pend.append(eng.segment_submit(None, thr, min_area_x2, _lib.PREPROC_MOBILENET, max_comps=cap,
                                           device_ptr=d_imgs[k & 1].data_ptr(), shape=(B, H, W), dtype=_lib.UBD_U8))
            if len(pend) == DEPTH:
                _, counts = eng.segment_wait(pend.pop(0))
        while pend:
            _, counts = eng.segment_wait(pend.pop(0))
        return counts

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local_rank)
    clk.__enter__()
    run_dev(args.warmup)
    eng.set_option("profile", 1)
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk.mark(True)
    e0.record(stream)
    counts = run_dev(args.steps)
    e1.record(stream)
    barrier()
    clk.mark(False)
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    dil_ms, dil_n = eng.stat("dilconv_ms"), eng.stat("dilconv_launches")
    stem_ms, stem_n = eng.stat("stem_ms"), eng.stat("stem_launches")
    ccl_ms, head_ms = eng.stat("ccl_ms"), eng.stat("head_ms")
    host_ms = [eng.stat(f"host_ms{i}") / args.steps for i in range(5)]
    eng.set_option("profile", 0)

    # the same loop with 16-bit map containers, reported beside the headline (never instead of it): "f16" = IEEE-half maps and
    # weights with fp32 accumulation - the 10-bit significand a tf32 MMA reads, at half the map traffic - and "bf16"
    other = {}
    if precision == "tf32" and not args.no_extra:
        for p2 in ("f16", "bf16"):
            e2 = Engine(device=local_rank, precision=p2, n_classes=n_classes)
            e2.set_weights(weights)
            e2.set_stream(stream.cuda_stream)
            run_dev(args.warmup, e2)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            run_dev(args.steps, e2)
            a1.record(stream)
            barrier()
            other[p2] = a0.elapsed_time(a1)
            e2.set_stream(None)
            del e2

    # end to end through the host-buffer C-ABI calls (what ModelRunner.predict / predict_stream do): pinned host
    # images in, mask + components out, every copy inside the timed region
    mask_h = [torch.empty((B, H // 4, W // 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(DEPTH)]

    def run_e2e(steps):
        counts_h = None
        if args.sync_api:
            for k in range(steps):
                comps_h = np.zeros(cap, _lib.COMPONENT_DTYPE)
                counts_h = np.zeros(B, np.int32)
                _lib.check(eng.handle, eng._lib.ubd_segment(eng.handle, _lib.ptr(h_imgs[k & 1]), _lib.UBD_U8, B, H, W, _lib.PREPROC_MOBILENET,
                                                             np.float32(thr), min_area_x2, _lib.ptr(mask_h[k & 1]), None, None,
                                                             _lib.ptr(comps_h), cap, _lib.ptr(counts_h)))
            return counts_h
        pend = []
        for k in range(steps):
            pend.append(eng.segment_submit(h_imgs[k & 1], thr, min_area_x2, _lib.PREPROC_MOBILENET, mask_out=mask_h[k % DEPTH], max_comps=cap))
            if len(pend) == DEPTH:
                _, counts_h = eng.segment_wait(pend.pop(0))
        while pend:
            _, counts_h = eng.segment_wait(pend.pop(0))
        return counts_h

    run_e2e(max(DEPTH + 1, args.warmup))            # every result slot has staged a host batch once (its buffers exist)
    barrier()
    import gc
    gc.collect()
    gc.disable()                                    # the timed region is ~50 ms of wall clock: no collector pauses inside
    clk.mark(True)
    t0 = time.perf_counter()
    counts_h = run_e2e(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk.mark(False)
    gc.enable()
    n_comp_last = int(counts_h.sum())
    clk.__exit__()

    # config E beside it: training step (forward + loss + backward + Adam) on 32 x 512x512 per GPU,
    # gradients all-reduced over NCCL when N > 1 (the only collective on the path)
    train_ms = float("nan")
    train_loss = float("nan")
    if not args.no_train:
        from ubdvss_b200 import synth
        from ubdvss_b200.net import Adam, B200Model, NetConfig
        from ubdvss_b200 import losses as ulosses
        eng.set_stream(None)
        tb = 32
        # pinned host batches (what a loader hands over); a tf32 handle runs the dilated layers' forward, dgrad and wgrad on
        # the tensor cores (UBD_TRAIN_PRECISION=fp32: the exact FP32-pipe step)
        train_prec = os.environ.get("UBD_TRAIN_PRECISION", "tf32")
        tx = torch.from_numpy(np.concatenate([synth.synth_images(8, 512, 512, seed=40 + rank)] * (tb // 8))).pin_memory().numpy()
        ty = torch.from_numpy(np.concatenate([synth.synth_targets(8, 128, 128, 0, seed=40 + rank)] * (tb // 8)).astype(np.int32)).pin_memory().numpy()
        model = B200Model(NetConfig(), device=local_rank, precision=train_prec,
                          weights=usynth.synth_weights(0, seed=1234, calibrated=True))
        model.compile(Adam(1e-3), loss=ulosses.get_loss(False))
        if dist is not None:
            model.set_distributed(True)
        for _ in range(2):
            model.train_on_batch(tx, ty, preprocessing="mobilenet_like")
        barrier()
        t0 = time.perf_counter()
        tsteps = max(2, min(args.steps, 20))
        for _ in range(tsteps):
            out = model.train_on_batch(tx, ty, preprocessing="mobilenet_like")
        torch.cuda.synchronize()
        train_ms = (time.perf_counter() - t0) * 1e3 / tsteps
        train_loss = out[0]

    t = torch.tensor([ms, e2e_s * 1e3, train_ms, other.get("f16", 0.0), other.get("bf16", 0.0)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, train_ms_max = float(t[0]), float(t[1]), float(t[2])
    other = {k: float(t[3 + i]) for i, k in enumerate(("f16", "bf16")) if k in other}
    total_images = B * world * args.steps
    value = total_images / (ms_max / 1e3)
    e2e_value = total_images / (e2e_ms_max / 1e3)

    if rank == 0:
        pk = peaks()
        q_px = (H // 4) * (W // 4)
        tensor_rate = 0.5 if precision in ("fp32", "tf32") else 1.0       # tf32 MMAs run at half the bf16 rate
        peak_burst, peak_sust = pk["bf16_tflops"] * tensor_rate, pk["bf16_tflops_sustained"] * tensor_rate
        # the net is compute-bound (AI ~ 965 FLOP/B), so the roofline of the dominant kernel is the tensor pipe.
        # Dominant kernel = the dilated 3x3 24->24 layer: ALGORITHMIC 2 x 5,184 FLOP per map pixel per layer x the map
        # pixels one launch processes / its average duration (CUDA events on the launching stream inside the library).
        imgs_per_launch = B * 6 * args.steps / max(dil_n, 1)
        avg_launch_s = dil_ms / 1e3 / max(dil_n, 1)
        flops_per_launch = 2.0 * DIL_MAC_PER_MAP_PX * q_px * imgs_per_launch
        achieved_tf = flops_per_launch / avg_launch_s / 1e12 if avg_launch_s > 0 else 0.0
        stem_flops_per_step = 2.0 * (33 + 792 + 792 / 4.0) * (H // 2) * (W // 2) * B      # L1 + L2 at half, L3 at quarter resolution
        # DRAM bytes of one step, measured in this run: a child process runs one step under
        # `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none` (caches left as in a real run)
        traffic, traffic_src, wasted, step_dram = None, None, None, None
        if world == 1 and not args.no_traffic:
            traffic, step_dram, traffic_src = measure_traffic(args, precision)
            if step_dram:
                wasted = step_dram / (ALGO_BYTES_PER_IMAGE_1024 * (H * W / 1048576.0) * B)
        step_flops = 2.0 * ALGO_MAC_PER_INPUT_PX * H * W * B
        roof = {"bound": "tensor", "kernel": "tc4::dilconv_col_kernel: dilated 3x3 conv 24->24 (L4-L9), one layer of one chunk per launch",
                "achieved": achieved_tf, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved_tf / peak_burst if peak_burst else None,
                "frac_of_sustained_peak": achieved_tf / peak_sust if peak_sust else None,
                "peak_source": f"{pk['source']} bf16 burst {pk['bf16_tflops']} TF/s (sustained {pk['bf16_tflops_sustained']})"
                               + (" x 0.5 (tf32 rate)" if tensor_rate == 0.5 else ""),
                "algorithmic_flops_per_launch": flops_per_launch, "avg_launch_us": avg_launch_s * 1e6, "launches": int(dil_n),
                "traffic": traffic, "traffic_source": traffic_src,
                "share_of_step": dil_ms / ms if ms else None,
                "stem": {"kernels": "tc4::dilconv_col_kernel<L1SRC> (L1+L2) + tc::dilconv_tc_kernel (L3)" if stem_n >= 2 * max(dil_n, 1) / 6 - 0.5
                                    else "stemf::stem_fused_kernel (L1+L2+L3)",
                         "ms_per_step": stem_ms / args.steps, "share_of_step": stem_ms / ms if ms else None,
                         "algorithmic_tflops": stem_flops_per_step / (stem_ms / args.steps / 1e3) / 1e12 if stem_ms else None,
                         "frac": stem_flops_per_step / (stem_ms / args.steps / 1e3) / 1e12 / peak_burst if stem_ms and peak_burst else None,
                         "note": "two kernels (L1+L2, L3); as ONE kernel function the dilated layer above has the largest total time per step"},
                "whole_step": {"algorithmic_tflops": step_flops * value / (B * world) / 1e12,
                               "tensor_frac": step_flops * value / (B * world) / 1e12 / peak_burst if peak_burst else None,
                               "hbm_frac": ALGO_BYTES_PER_IMAGE_1024 * (H * W / 1048576.0) * (value / world) / (pk["hbm_gbs"] * 1e9),
                               "algorithmic_bytes_per_image": ALGO_BYTES_PER_IMAGE_1024 * (H * W / 1048576.0),
                               "dram_bytes_per_step": step_dram, "wasted_traffic_ratio": wasted},
                "cc": {"ms_per_step": ccl_ms / args.steps, "algorithmic_bytes_per_step": 5 * q_px * B,
                       "hbm_frac": 5 * q_px * B / (ccl_ms / args.steps / 1e3) / (pk["hbm_gbs"] * 1e9) if ccl_ms else None},
                "stage_ms_per_step": {"stem": stem_ms / args.steps, "dilated": dil_ms / args.steps,
                                      "head": head_ms / args.steps, "ccl": ccl_ms / args.steps},
                "host_ms_per_step": dict(zip(["enqueue_forward", "enqueue_cc", "wait_counts", "read_records", "boxes"], host_ms))}
        line = {"metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16", "f16": "f16"}[precision],
                "data": "synthetic", "config": cfg, "clocks": clk.summary(),
                "api": "ubd_segment_dev / ubd_segment" if args.sync_api else f"ubd_segment_submit[_dev] + ubd_segment_wait ({DEPTH} batches in flight)",
                "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": int(imgs[0].nbytes),
                        "d2h_bytes_per_step": int(mask_h[0].nbytes + 4 * B + 44 * n_comp_last), "ms_per_step": e2e_ms_max / args.steps},
                "cpu_affinity": numa, "gpu_launches": int(launches), "roofline": roof, "components_last_step": int(counts.sum()),
                "components_per_image": float(counts.sum()) / B}
        if other:
            line["other_containers"] = {k: {"value": total_images / (v / 1e3), "unit": "images/sec", "ms_per_step": v / args.steps,
                                            "note": {"f16": "IEEE-half maps + weights, fp32 accumulate: tf32's 10-bit significand, half the map bytes; saturates at 65504",
                                                     "bf16": "bf16 maps + weights, fp32 accumulate (configs[3] precision)"}[k]}
                                        for k, v in other.items()}
        if not args.no_train:
            line["train_step"] = {"config": f"configs[4]: forward + losses.py loss + backward + Adam, batch 32 of 512x512 per GPU, {train_prec}"
                                            + (", gradient all-reduce over NCCL inside libubd (ubd_allreduce_grads)" if world > 1 else ""),
                                  "ms_per_step": train_ms_max, "images_per_sec": 32 * world / (train_ms_max / 1e3),
                                  "loss": train_loss}
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core again
            v, step_s, cores = time_cpu(weights, imgs[0][:args.cpu_sample], thr, n_classes, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "images/sec", "cores": cores, "kind": "port",
                                    "sample": f"3 steps x {args.cpu_sample} images of {H}x{W} (torch-CPU restatement + cv2)"}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
