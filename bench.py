#!/usr/bin/env python
"""bench.py — VarDCT encode throughput (MP/s) of the B200 path, with the roofline of its
dominant kernel and the CPU baseline beside it.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference ...                   (CPU arm: the oracle on the host cores)

A step = one batch of `--batch` synthetic images of the workload per rank, `--pipelines` of them in flight (the
reference keeps 6 workers busy the same way, benchmark-jpegxl/src/config.rs:22).  The default workload is the
configuration BASELINE.json's metric is quoted on, configs[4]: 1920x1080 RGB8, combined.diff (both proposals' hooks),
effort 7, distance 0.5 .. 3.0 round-robin by global image index, image i on rank i mod N.  `value` is measured with the
images resident in HBM (jxlb200_encode_batch_device, CUDA events on the encoder's streams);
`e2e` goes through jxlb200_encode with HOST buffers (pinned staging + H2D + kernels + D2H
of the codestream inside the timed region).  Ranks shard by image (no data-path collective):
weak scaling; NCCL only gathers the per-rank times.
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per pipeline stream (before CUDA init)
sys.path.insert(0, ROOT)
PKG = "jpeg-xl-lossy-image-compression-thesis_b200"

WORKLOADS = {
    # name: (width, height, distance, effort, proposal, flags)
    "4k_dct8_d1": (3840, 2160, 1.0, 7, 0, 1),
    "4k_full_d1": (3840, 2160, 1.0, 7, 0, 0),
    "8k_partitioning": (7680, 4320, 1.0, 7, 1, 0),
    "1080p_combined": (1920, 1080, 1.0, 7, 3, 0),
    # BASELINE configs[4]: 1080p, combined proposal, distance 0.5 .. 3.0 round-robin by GLOBAL image index, image i on
    # rank i mod N (sharding.distance_for_image / shard_indices); the per-image distance list replaces `distance`
    "1080p_combined_sweep": (1920, 1080, None, 7, 3, 0),
    "512_d1": (512, 512, 1.0, 7, 0, 0),
}
# algorithmic bytes per pixel of each pipeline stage (DESIGN.md "Kernels")
STAGE_BYTES_PER_PX = {"xyb": 15.0, "aq": 12.1, "homog": 12.2, "coeff": 18.3}
# DRAM bytes of one launch from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum) of the DCT8 frame's
# transform + quantise kernel at 3840x2160 (profiles/r01f_dct8_v4_recon_full.txt: 100.1 MB read + 20.3 MB written);
# the search workloads' coefficient stage is a dozen launches (one per strategy), no single capture applies: null
TRAFFIC_BYTES_K7_DCT8_4K = 120.4e6
PEAK_WARP_INST_PER_S = 148 * 4 * 1.965e9   # issue slots of the GPU: 148 SMs x 4 schedulers x max SM clock
ACS_WARP_INST_PER_PX = 106.0   # search kernels, warp instructions per pixel at 1080p combined d = 1 (profiles/r02m_launches_1080p.csv: 219.8 M per frame)
STAGE_INDEX = {"h2d": 0, "xyb": 1, "aq": 2, "homog": 3, "acs": 4, "coeff": 5, "tokenize": 6, "histo": 7, "ans": 8,
               "dc": 9, "assemble": 10, "d2h": 11}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_throughput(workload, budget_images, threads):
    """Times the CPU oracle (oracle/_build/libjxo.so — the repo's scalar restatement; libjxl itself is
    not available offline) on `budget_images` images of the workload using `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    pkg = importlib.import_module(PKG)
    w, h, dist, effort, proposal, flags = WORKLOADS[workload]
    sweep = dist is None
    ora = oracle_lib.load(rebuild=not os.path.exists(oracle_lib.SO))
    imgs = [pkg.synth_image(w, h, 1000 + i) for i in range(min(budget_images, 2))]
    done = []

    def work(k):
        f = ora.encode(imgs[k % len(imgs)], pkg.distance_for_image(k) if sweep else dist, effort, proposal, flags)
        assert f.error == "", f.error
        done.append(len(f.dump("codestream")))
        f.close()

    t0 = time.perf_counter()
    pending = list(range(budget_images))
    running = []
    while pending or running:
        while pending and len(running) < threads:
            th = threading.Thread(target=work, args=(pending.pop(),))
            th.start()
            running.append(th)
        running[0].join()
        running.pop(0)
    dt = time.perf_counter() - t0
    return budget_images * w * h / 1e6 / dt, dt


def find_cjxl():
    """The reference's encoder binary, if one was supplied ($JXLB200_CJXL, baseline/_ref/bin/cjxl, PATH): SURVEY 8d.
    None in this image (libjxl is cloned from the network inside the reference's Docker build, Dockerfile:40)."""
    import shutil
    for cand in (os.environ.get("JXLB200_CJXL"), os.path.join(ROOT, "baseline", "_ref", "bin", "cjxl"), shutil.which("cjxl")):
        if cand and os.path.isfile(cand) and os.access(cand, os.X_OK):
            return cand
    return None


def cjxl_throughput(cjxl, workload, n_images, threads):
    """`cjxl in.ppm out.jxl --distance=D --effort=E` (the reference's call, docker_manager.rs:125-136) on synthetic
    PPMs, libjxl's own thread pool on `threads` threads, one image after the other."""
    import tempfile
    pkg = importlib.import_module(PKG)
    w, h, dist, effort, proposal, flags = WORKLOADS[workload]
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for i in range(min(n_images, 2)):
            path = os.path.join(tmp, f"in{i}.ppm")
            with open(path, "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (w, h))
                f.write(pkg.synth_image(w, h, 1000 + i).tobytes())
            paths.append(path)
        t0 = time.perf_counter()
        for i in range(n_images):
            d_i = importlib.import_module(PKG).distance_for_image(i) if dist is None else dist
            subprocess.run([cjxl, paths[i % len(paths)], os.path.join(tmp, "out.jxl"), f"--distance={d_i}", f"--effort={effort}",
                            f"--num_threads={threads}"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
    return n_images * w * h / 1e6 / dt, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference's own
    encoder (libjxl behind `cjxl`, docker_manager.rs:136) cannot be built or installed offline, so this
    arm times the oracle port with one image per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w, h, dist, effort, proposal, flags = WORKLOADS[args.workload]
    per_step = max(1, cores)            # one oracle encode per host thread, every core of the box
    cjxl = find_cjxl()
    vals = []
    for i in range(args.warmup + args.steps):
        if cjxl:
            v, dt = cjxl_throughput(cjxl, args.workload, 4, cores)
        else:
            v, dt = cpu_oracle_throughput(args.workload, per_step, per_step)
        if i >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    ms = statistics.mean(dt for _, dt in vals) * 1e3
    line = {
        "impl": "reference", "metric": "vardct_encode_throughput", "value": value, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "width": w, "height": h,
                   "distance": "0.5..3.0 round-robin by global image index" if dist is None else dist, "effort": effort,
                   "proposal": proposal, "flags": flags},
        "cpu_baseline": ({"value": value, "unit": "MP/s", "cores": cores, "kind": "reference",
                          "sample": f"4 images of {w}x{h} per step through {cjxl} --num_threads={cores}"} if cjxl else
                         {"value": value, "unit": "MP/s", "cores": per_step, "kind": "port",
                          "sample": f"{per_step} images of {w}x{h} per step, one oracle encode per host thread "
                                    "(own CPU restatement; no cjxl binary found: libjxl is not installable offline)"}),
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p_combined_sweep", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-k7", action="store_true", help="skip the 4K DCT8 sub-record (roofline.k7_dct8)")
    ap.add_argument("--batch", type=int, default=256, help="images per rank per step (the pipeline's fill and drain, a few ms "
                    "of rANS latency per image, amortise over the step)")
    ap.add_argument("--pipelines", type=int, default=32, help="images in flight per rank (CUDA streams)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module(PKG)
    # one process per GPU: run on (and allocate the pinned input buffers from) the GPU's own NUMA node
    numa = pkg.bind_to_gpu_numa_node(local_rank) if os.environ.get("JXLB200_NUMA_BIND", "1") != "0" else {"numa_node": "off"}
    w, h, distance, effort, proposal, flags = WORKLOADS[args.workload]
    mp = w * h / 1e6
    B = args.batch                               # images per rank per step, args.pipelines of them in flight
    n_distinct = min(B, 8)                       # distinct synthetic images per rank (the batch cycles through them)
    imgs = [pkg.synth_image(w, h, rank * 64 + i) for i in range(n_distinct)]
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]          # e2e inputs: page-locked host memory
    if distance is None:    # config 5: this rank's images are the global indices rank, rank + world, ...
        distance = [pkg.distance_for_image(rank + world * i) for i in range(B)]
    h_imgs = [pinned[i % n_distinct].numpy() for i in range(B)]
    d_imgs = [p.cuda() for p in pinned]                                  # value inputs: resident in HBM
    d_ptrs = [d_imgs[i % n_distinct].data_ptr() for i in range(B)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    enc = pkg.Encoder(local_rank)
    enc.set_pipelines(args.pipelines)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM; CUDA-event time of each batch (all pipeline streams) ----
    for i in range(args.warmup):
        enc.encode_batch_device(d_ptrs, w, h, 3 * w, distance, effort, proposal, flags)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, launches = [], 0
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 255)                     # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        sts, ms = enc.encode_batch_device(d_ptrs, w, h, 3 * w, distance, effort, proposal, flags)
        dev_ms.append(ms)
        launches += sum(s.kernel_launches for s in sts)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = float(sum(dev_ms))
    last_stats = sts[-1]

    # ---- per-kernel times for the roofline: one image alone on the GPU (no overlap between streams) ----
    stage_ms = []
    for i in range(max(3, min(args.steps, 10))):
        flush.fill_(i & 255)
        torch.cuda.synchronize()
        st1 = enc.encode_device(d_ptrs[i % n_distinct], w, h, 3 * w, distance[i % len(distance)] if isinstance(distance, list) else distance,
                                effort, proposal, flags)
        stage_ms.append(list(st1.stage_ms) + [st1.total_ms])

    # ---- e2e: pinned host buffers through jxlb200_encode_batch, wall clock around the calls ----
    for i in range(args.warmup):
        enc.encode_batch(h_imgs, distance, effort, proposal, flags)
    barrier()
    t0 = time.perf_counter()
    out_bytes = 0
    for i in range(args.steps):
        data, st2 = enc.encode_batch(h_imgs, distance, effort, proposal, flags)
        out_bytes += sum(len(d) for d in data)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- single image through the reference-facing call (host buffer in, codestream out), as the harness calls cjxl
    single_ms = []
    for i in range(5):
        t1 = time.perf_counter()
        enc.encode(h_imgs[i % n_distinct], distance[i % len(distance)] if isinstance(distance, list) else distance, effort, proposal, flags)
        single_ms.append((time.perf_counter() - t1) * 1e3)
    # ---- sub-record: the DCT8-only transform + quantise kernel (K7) and XYB (K1) at 3840x2160, same run
    k7 = None
    if rank == 0 and not args.no_k7:
        img4k = torch.from_numpy(pkg.synth_image(3840, 2160, 7)).cuda()
        st4 = []
        for i in range(6):
            flush.fill_(i & 255)
            torch.cuda.synchronize()
            s4 = enc.encode_device(img4k.data_ptr(), 3840, 2160, 3 * 3840, 1.0, 7, 0, 1)
            if i:
                st4.append(list(s4.stage_ms))
        m4 = np.mean(np.array(st4), axis=0)
        px4 = 3840 * 2160
        pk, _ = measured_peak_gbs()
        k7 = {"workload": "4k_dct8_d1 (BASELINE configs[1]), single-image encodes",
              "kernel": "k_dct8_quant_v4", "kernel_ms": float(m4[STAGE_INDEX["coeff"]]),
              "achieved": 18.3 * px4 / (float(m4[STAGE_INDEX["coeff"]]) / 1e3) / 1e9, "peak": pk, "unit": "GB/s",
              "frac": 18.3 * px4 / (float(m4[STAGE_INDEX["coeff"]]) / 1e3) / 1e9 / pk, "traffic": TRAFFIC_BYTES_K7_DCT8_4K,
              "xyb": {"kernel": "k_rgb8_to_xyb", "kernel_ms": float(m4[STAGE_INDEX["xyb"]]),
                      "frac": 15.0 * px4 / (float(m4[STAGE_INDEX["xyb"]]) / 1e3) / 1e9 / pk}}
        del img4k

    # per-rank statistics -> job totals (SUM of counters, MAX of times): the only communication of the job
    job = pkg.gather_stats(pkg.ShardStats(images=B * args.steps, pixels=B * args.steps * w * h,
                                          codestream_bytes=out_bytes, device_ms=total_ms, kernel_launches=launches),
                           device="cuda")
    job_e2e = pkg.gather_stats(pkg.ShardStats(device_ms=e2e_s * 1e3), device="cuda")
    total_ms, e2e_s, launches, out_bytes = job.max_ms, job_e2e.max_ms / 1e3, job.kernel_launches, job.codestream_bytes // world

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * B * mp / (ms_per_step / 1e3)
        e2e_value = world * B * mp * args.steps / e2e_s
        mean_stage = np.mean(np.array(stage_ms), axis=0)
        peak, peak_kind = measured_peak_gbs()
        search = not (flags & 1) and effort >= 5
        # the HBM-bound stage the north star names: transform + quantise (18.3 B/px: 12 in, 6 out, 0.3 side data).  On
        # DCT8-only frames that is ONE kernel (k_dct8_quant_v4); on search workloads it is the strategy-binned stage
        # (k_coeff_lists + k_coeff8_lanes + k_coeff8_special + k_coeffsq_all<16|32|64>, one launch per size class), timed as a stage.
        dom = "coeff"
        dom_ms = float(mean_stage[STAGE_INDEX[dom]])
        achieved = STAGE_BYTES_PER_PX[dom] * w * h / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        per_stage = {}
        for sname, bpp_alg in STAGE_BYTES_PER_PX.items():
            if sname == "homog" and proposal == 0:
                continue                                    # the map is only computed for the proposals (a memset otherwise)
            t_ms = float(mean_stage[STAGE_INDEX[sname]])
            per_stage[sname] = {"ms": t_ms, "GB/s": bpp_alg * w * h / (t_ms / 1e3) / 1e9 if t_ms > 0 else 0.0,
                                "frac": (bpp_alg * w * h / (t_ms / 1e3) / 1e9 / peak) if t_ms > 0 else 0.0}
        acs_ms = float(mean_stage[STAGE_INDEX["acs"]])
        roof = {"bound": "hbm",
                "kernel": ("coefficient stage of the search path: k_coeff_lists + k_coeff8_lanes + k_coeff8_special + "
                           "k_coeffsq_all<16|32|64> (transform + quantise, one launch per size class)") if search
                          else "k_dct8_quant_v4 (transform + quantise)",
                "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if search else (TRAFFIC_BYTES_K7_DCT8_4K if (w, h) == (3840, 2160) else None),
                "algorithmic_bytes_per_launch": STAGE_BYTES_PER_PX[dom] * w * h, "kernel_ms": dom_ms,
                "hbm_stages": per_stage,
                "single_image_stage_ms": {k: float(mean_stage[v]) for k, v in STAGE_INDEX.items()},
                "single_image_total_ms": float(mean_stage[-1]),
                "nominal_peak": {"GB/s": 8000.0, "frac": achieved / 8000.0},
                "entropy_stage": {"kernel": "k_ans_groups (one warp per 256x256 AC group, serial rANS chain)",
                                  "ms": float(mean_stage[STAGE_INDEX["ans"]]), "tokens": int(last_stats.num_tokens),
                                  "Mtokens_per_s": (last_stats.num_tokens / 1e6) / (float(mean_stage[STAGE_INDEX["ans"]]) / 1e3)
                                  if mean_stage[STAGE_INDEX["ans"]] > 0 else 0.0,
                                  "bound": "dependency latency of the longest group's chain, not bandwidth (profiles/)"},
                "note": "per-kernel times from single-image encodes (one stream)"}
        if search:
            # the stage that dominates the step by time is not a bandwidth stage: the AC-strategy search evaluates ~20
            # candidate transforms per pixel (FP32-issue bound); reported against the GPU's issue slots
            inst = ACS_WARP_INST_PER_PX * w * h
            roof["dominant_by_time"] = {
                "kernel": "AC-strategy search: k_acs_evalsq<8|16|32|64> + k_acs_eval8s + k_acs_decide", "bound": "fp32 issue",
                "ms": acs_ms, "warp_instructions": inst, "achieved_Ginst_per_s": inst / (acs_ms / 1e3) / 1e9 if acs_ms > 0 else 0.0,
                "peak_Ginst_per_s": PEAK_WARP_INST_PER_S / 1e9,
                "frac": inst / (acs_ms / 1e3) / PEAK_WARP_INST_PER_S if acs_ms > 0 else 0.0,
                "source": "warp instructions per pixel from ncu smsp__inst_executed.sum (profiles/r02m_launches_1080p.csv)"}
        if k7 is not None:
            roof["k7_dct8"] = k7
        line = {
            "metric": "vardct_encode_throughput", "value": value, "unit": "MP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "width": w, "height": h,
                       "distance": "0.5..3.0 round-robin by global image index" if isinstance(distance, list) else distance,
                       "effort": effort,
                       "proposal": proposal, "flags": flags, "images_per_rank_per_step": B,
                       "pipelines_per_rank": args.pipelines,
                       "l2": "flushed between timed iterations (256 MiB fill, untimed); distinct inputs per step "
                             f"{n_distinct * 3 * w * h >> 20} MiB",
                       "parallelism": f"image-sharded x{world}, no data-path collective",
                       "host_placement": f"rank 0 bound to NUMA node {numa.get('numa_node')} ({numa.get('cpus')} cpus)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": world * B * 3 * w * h,
                    "d2h_bytes_per_step": world * out_bytes // max(1, args.steps), "ms_per_step": e2e_s * 1e3 / args.steps,
                    "h2d_GBps_per_gpu": B * 3 * w * h / (e2e_s / args.steps) / 1e9},
            "gpu_launches": launches,
            "roofline": roof,
            "single_image_e2e_ms": float(np.median(single_ms)),
            "bpp": last_stats.bpp, "codestream_bytes": last_stats.codestream_bytes,
            "wall_s_value_loop": wall_s,
        }
        if not args.no_cpu_baseline:
            # rank 0 times the CPU implementation on every host core (one oracle encode per thread), at every N
            cores = os.cpu_count() or 1
            n_img = max(cores, 2)
            v, dt = cpu_oracle_throughput(args.workload, n_img, cores)
            line["cpu_baseline"] = {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_img} images of {w}x{h}, one scalar oracle encode per host thread on {cores} "
                                              f"threads, {dt:.1f} s (own CPU restatement: no libjxl binary exists offline)"}
        print(json.dumps(line), flush=True)
    enc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
