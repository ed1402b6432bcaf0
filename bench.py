#!/usr/bin/env python
"""bench.py -- RAW->bitstream forward throughput of the B200 path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the reference's algorithm on the host CPU cores)

One "step" = one full pass of raw_compression_tcm_final over one synthetic 4 x T x T packed-Bayer tile
per GPU (BASELINE config[1]: T = 2048): analysis, hyper-prior, 5-slice entropy parameters + likelihoods,
synthesis to x_hat, AND the integer range coder producing the y/z bitstreams (forward(x, emit_strings=True)).
Prints ONE JSON line (rank 0).  1 MP = 1e6 sensor photosites; a tile holds 4*T*T of them.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "megapixels_per_s_raw_to_bitstream_forward"
FLOP_PER_PACKED_POS = 2.764e6      # raw_compression_tcm_final.forward, 2*MAC per packed position
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch, (engine, T) -> bytes, from the committed ncu captures
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel, from the `ncu --set full` captures summarised in
# profiles/r2_conv_tail_ncu.md (gpurun_out/r2_t17_tail*_raw.csv): key = (engine of the launch, tile size)
NCU_TRAFFIC = {("fp16", 2048): 1.074746e9 + 1.032225e9, ("bf16x3", 2048): 2.171822e9 + 2.107215e9}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        d["_src"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def _nvml(self):
        """In-process NVML sampling (no fork of a process that maps tens of GB of device memory)."""
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[n] else "Not Active"
                                                   for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            self._halt.wait(0.05)

    def run(self):
        try:
            return self._nvml()
        except Exception:
            pass
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1] else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------- CPU reference leg
def _cpu_arm(T):
    """The reference's own implementation of the path on the host: the UNMODIFIED reference modules when oracle/_ref is staged
    (byte-compiled by oracle/stage_ref.py in the authoring container; CompressAI, absent offline, is bound to the restatement in
    oracle/cai.py), else the functional port oracle/refpath.py.  Returns (step, kind, note); step() -> (bitstream, symbols)."""
    import numpy as np
    import torch

    from oracle import ref_import, refpath
    from realcamnet_b200 import raw2bit  # parameter names/shapes only; nothing of it runs in this leg
    from realcamnet_b200 import synthetic

    x = synthetic.make_inputs(T, seed=1234)
    gc = refpath._gc()
    nsl = 5

    def encode(para):
        sym = torch.round(para["y"] - para["means"]).to(torch.int32)
        idx = gc.build_indexes(para["scales"])
        C = sym.shape[1]
        # slice-major order, as compress() flattens them (raw2bit.py:1943-1944)
        s = np.concatenate([sym[:, i * C // nsl:(i + 1) * C // nsl].reshape(-1).numpy() for i in range(nsl)])
        i_ = np.concatenate([idx[:, i * C // nsl:(i + 1) * C // nsl].reshape(-1).numpy() for i in range(nsl)])
        return refpath.encode_stream(s, i_, gc), sym

    if ref_import.reference_available():
        ref = ref_import.import_reference()
        torch.set_grad_enabled(False)
        m = ref.raw2bit.raw_compression_tcm_final().eval()
        synthetic.fill_(m, seed=0)

        def step():
            return encode(m(x)["para"])
        return step, "reference", ("unmodified reference modules (models/raw2bit.py raw_compression_tcm_final.forward, eval) from "
                                   f"{'/root/reference' if ref_import.reference_root() == ref_import.REFERENCE_ROOT else 'oracle/_ref (byte-compiled)'}"
                                   " + range coding of its symbols; CompressAI layers/entropy models = oracle/cai.py restatement")
    pm = raw2bit.raw_compression_tcm_final()
    synthetic.fill_(pm, seed=0)
    sd = {k: v.detach().clone() for k, v in pm.state_dict().items()}

    def step():
        return encode(refpath.final_forward(sd, x)["para"])
    return step, "port", "oracle restatement (oracle/refpath.py of raw2bit.py:1766-1855) + range coding of its symbols"


def run_cpu_reference(T, steps, warmup, budget_s=150.0):
    """Times the CPU arm on a 4xTxT tile.  torch's CPU kernels do not scale to 100+ threads on these layer sizes, so the thread
    count is probed (all cores, 32, 16) on a small 4x512x512 tile first and the fastest is used and reported.  At most `steps`
    timed passes, fewer when they would exceed the wall budget (each pass at T=2048 is tens of seconds)."""
    import torch

    cores = os.cpu_count() or 1
    probe, _, _ = _cpu_arm(512)
    best, best_t = cores, None
    probe()                                # untimed: imports, allocator and thread-pool warm-up
    for th in sorted({cores, min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(th)
        t0 = time.perf_counter()
        probe()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    del probe
    torch.set_num_threads(best)
    step, kind, note = _cpu_arm(T)
    sym = None
    for _ in range(max(1, min(warmup, 1))):
        _, sym = step()
    done, t0 = 0, time.perf_counter()
    while done < max(1, steps):
        step()
        done += 1
        el = time.perf_counter() - t0
        if el + el / done > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    mp = 4.0 * T * T / 1e6
    return {"value": mp / dt, "dt": dt, "threads": best, "steps_timed": done, "kind": kind, "note": note, "sym": sym}


def workload_name(T):
    """the same string in both arms' `config.workload` (the driver compares the arms' configurations)"""
    return (f"raw_compression_tcm_final.forward + range coder on one 4x{T}x{T} packed-Bayer tile per GPU (BASELINE config[1]), "
            "random-init weights (name-keyed, seed 0)")


# --------------------------------------------------------------------------------------------- BASELINE config 4
def measure_frame4k(model, dev, rank, world, height=2160, width=3840, tile=512, reps=2):
    """Frames/s of the tiled-frame path on the ranks of this run (wall time incl. all host work, max over ranks)."""
    import hashlib

    import torch
    import torch.distributed as dist

    from realcamnet_b200 import dist as rdist
    from realcamnet_b200 import tiler

    model.enable_cuda_graphs(True)
    fr = torch.rand(4, height, width, generator=torch.Generator().manual_seed(99)).pin_memory() if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    blob = rdist.compress_frame_distributed(model, fr, height, width, tile, dev)      # warm-up: graph capture of the batch shape
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        blob2 = rdist.compress_frame_distributed(model, fr, height, width, tile, dev)
    barrier()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return None
    ny, nx = tiler.tile_grid(height, width, tile)
    mp = 4.0 * height * width / 1e6
    from realcamnet_b200 import frame as rframe
    return {"workload": f"BASELINE configs[3]: packed 4x{height}x{width} RAW frame -> {ny * nx} tiles of 4x{tile}x{tile}, tile t -> rank t mod {world}, "
                        "compress() per tile (batched, host coder overlapped), NCCL scatter / bitstream gather, RCNB container on rank 0",
            "value": mp / (ms / 1e3), "unit": "MP/s", "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "n_gpus": world, "tiles": ny * nx,
            "tiles_per_rank": len(rdist.my_tiles(ny * nx, 0, world)), "batch": rframe.batch_size_for(len(rdist.my_tiles(ny * nx, 0, world))),
            "container_bytes": len(blob), "container_sha256": hashlib.sha256(blob).hexdigest()[:16], "deterministic": blob2 == blob,
            "timing": "wall clock around whole frames incl. H2D of the frame, tiling, scatter, host range coder and gather; max over ranks"}


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tile", type=int, default=2048, help="packed tile side T (BASELINE config[1]: 2048)")
    ap.add_argument("--cpu-tile", type=int, default=0,
                    help="tile side of the CPU sample: default = --tile for --impl reference (like for like), 1024 for the cpu_baseline leg "
                         "of our own arm (~5 s per pass on 16 cores)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="wall budget (s) of the timed CPU passes of --impl reference")
    ap.add_argument("--tail-engine", default=os.environ.get("RCN_TAIL_ENGINE", "fp16"), choices=["fp16", "bf16x3", "bf16"],
                    help="per-stage precision policy: engine of the post-quantisation full-resolution synthesis tail when --engine is "
                         "bf16x3 (fp16 = one fp16 MMA pass, within the 1e-3 bar on x_hat; bf16x3 = policy off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frame", action="store_true", help="skip the secondary frame workload (BASELINE config 4)")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--engine", default=os.environ.get("RCN_CONV_ENGINE", "bf16x3"), choices=["fp32", "bf16x3", "bf16", "fp16"],
                    help="conv engine: fp32 = CUDA-core exact; bf16x3 = tcgen05 with hi/lo split operands (parity grade); "
                         "bf16 = tcgen05 single pass (fast mode, outside the 1e-3 parity bar)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        Tc = args.cpu_tile or args.tile
        r = run_cpu_reference(Tc, max(1, args.steps), max(0, min(args.warmup, 1)), args.cpu_budget)
        v, dt = r["value"], r["dt"]
        sample = (f"{r['note']}; one 4x{Tc}x{Tc} tile per step, torch CPU fp32, {r['threads']} threads (best of all-cores/32/16 probed "
                  f"on a 4x512x512 tile), 1 warm-up + {r['steps_timed']} timed passes ({dt:.1f} s each; --steps {args.steps} capped by a "
                  f"{args.cpu_budget:.0f} s wall budget); host has {os.cpu_count()} logical cores")
        line = {"metric": METRIC, "value": v, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "steps_timed": r["steps_timed"], "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload_name(Tc), "tile": Tc, "tiles_per_gpu": 1},
                "cpu_baseline": {"value": v, "unit": "MP/s", "cores": r["threads"], "kind": r["kind"], "sample": sample},
                "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from realcamnet_b200 import synthetic as inputs, synthetic as weights    # seeded inputs / name-keyed weights (not the oracle)
    from realcamnet_b200 import ops, raw2bit

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = args.tile
    ops.set_engine(args.engine)
    raw2bit.raw_compression_tcm_final.tail_engine = args.tail_engine
    tail = args.tail_engine if (args.engine == "bf16x3" and args.tail_engine != "bf16x3") else None   # policy active?
    model = raw2bit.raw_compression_tcm_final()
    weights.fill_(model, seed=0)
    model = model.to(dev).eval()
    model.update()
    model.enable_cuda_graphs(not args.no_graphs)
    x_host = [t.pin_memory() for t in inputs.make_inputs(T, seed=1234 + rank)]
    x_dev = [t.to(dev) for t in x_host]
    mp_tile = 4.0 * T * T / 1e6

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return model(x_dev, emit_strings=True)

    # End to end: every step copies ITS inputs from pinned host memory and ends with the bitstream bytes on the host.  The copy of step
    # k + 1's inputs is issued on a copy stream before step k's forward, so it overlaps with compute (an input pipeline around the
    # public call model(x, emit_strings=True)); K timed steps still contain K host->device copies (the one primed before the
    # timed region replaces the one issued for the step after the last).
    copy_stream = torch.cuda.Stream(device=dev)
    prefetched = []

    def prefetch():
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            xs = [t.to(dev, non_blocking=True) for t in x_host]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        for t in xs:
            t.record_stream(cur)
        prefetched.append((xs, ev))

    def step_e2e():
        if not prefetched:
            prefetch()
        xs, ev = prefetched.pop(0)
        torch.cuda.current_stream(dev).wait_event(ev)
        prefetch()                                 # next step's inputs travel while this step computes
        out = model(xs, emit_strings=True)
        return out["strings"]                      # bitstream bytes live on the host

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ops.launch_count()
    ms, out = timed(step_resident, args.steps)
    launches = ops.launch_count() - n0
    clocks = sampler.finish()
    value = world * mp_tile * args.steps / (ms / 1e3)
    step_e2e()
    ms_e2e, strings = timed(step_e2e, args.steps)
    e2e = world * mp_tile * args.steps / (ms_e2e / 1e3)
    nbytes = len(strings[0][0]) + sum(len(s) for s in strings[1])
    nsym = 320 * (T // 16) ** 2 + 192 * (T // 64) ** 2
    h2d = sum(t.numel() * 4 for t in x_host)
    # device->host per step: the coder front end's packed (start<<16|freq-1) and bypass words (int32 each) + escape flags (u8) per y
    # symbol, int32 symbols of z (tcm.py _begin_host_copy).  x_hat (3 x 2T x 2T fp32) stays on the device: the metric is
    # RAW -> bitstream.
    d2h = (4 + 4 + 1) * 320 * (T // 16) ** 2 + 4 * 192 * (T // 64) ** 2

    # ---- decode side: decompress(strings, shape) -> x_hat, host range decoder interleaved with the slice loop
    model.enable_cuda_graphs(False)
    shape = torch.Size([T // 64, T // 64])
    model.decompress(strings, shape)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ndec = 2
    for _ in range(ndec):
        model.decompress(strings, shape)
    torch.cuda.synchronize()
    decode_ms = (time.perf_counter() - t0) * 1e3 / ndec

    # ---- BASELINE config 4 on the same ranks: one 3840x2160 (packed 4 x 2160 x 3840) frame, 40 tiles of 4x512x512, tile t -> rank
    # t mod N, compress() per tile in equal batches with the host coder overlapped, scatter / gather over NCCL, one RCNB container
    frame4k = None
    if not args.no_frame:
        try:
            frame4k = measure_frame4k(model, dev, rank, world)
        except Exception as e:      # a secondary workload must never cost the bench line
            frame4k = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---- roofline of the dominant kernel: the full-resolution 128->128 3x3 conv + LeakyReLU of the g_s tail (ResidualBlock.conv1,
    # raw2bit.py:1681), launched exactly as the step launches it: the engine the precision policy gives the tail, operand planes
    # in, operand planes out (its fp32 result is never materialised).
    peaks = load_peaks()
    pc = ops.pack(model.g_s[10].conv1)
    a = torch.randn(1, T, T, 128, device=dev)
    kflop = 2.0 * 9 * 128 * 128 * T * T
    reps = 5
    keng = tail or args.engine
    if keng == "fp32":
        o = torch.empty_like(a)
        fn = lambda: ops.conv2d(a, pc, out=o, act=ops.ACT_LRELU, slope=0.01, engine="fp32")
        kname = "conv2d_kernel<128> (fp32 FFMA implicit GEMM)"
        peak, peak_note = float(peaks.get("bf16_tflops", 1590.0)), "burst bf16 (kernel timed alone); fp32 FFMA peak is ~75 TFLOP/s"
        work, kbytes = kflop, float(T) * T * 128 * 8
    else:
        passes = {"bf16x3": 3, "bf16": 1, "fp16": 1}[keng]
        with ops.engine_scope(keng):
            sp = ops.split_operand(a, pc.cp)

        def fn():
            with ops.engine_scope(keng):
                ops.conv2d(a, pc, presplit=sp, act=ops.ACT_LRELU, slope=0.01, emit_split=True, keep_fp32=False)
        kname = (f"conv_tc_kernel (tcgen05.mma + TMA, {passes} {'fp16' if keng == 'fp16' else 'bf16'} pass{'es' if passes > 1 else ''}"
                 ", operand planes in / out)")
        peak, peak_note = float(peaks.get("bf16_tflops", 1590.0)), "burst bf16 (kernel timed alone)"
        work = kflop * passes          # tensor-pipe FLOPs actually issued (hi*hi + lo*hi + hi*lo for bf16x3)
        kbytes = float(T) * T * 128 * 2 * (2 if passes == 1 else 4)   # 16-bit planes in + out (hi only | hi + lo)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / reps
    ach = kflop / (kms / 1e3) / 1e12        # ALGORITHMIC FLOPs of the convolution (2*k*k*Cin*Cout per output pixel) / time
    issued = work / (kms / 1e3) / 1e12      # what the tensor pipe executed (3 bf16 MMAs per product in the bf16x3 engine)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of THIS launch at T=2048
    # (profiles/r2_conv_tail_ncu_full.md); other engines / tile sizes were not captured.
    traffic = NCU_TRAFFIC.get((keng, T))
    roofline = {"kernel": kname + " -- 3x3 128->128 @ full res (g_s tail)", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "ms_per_launch": kms, "peak_source": peaks["_src"] + ", " + peak_note,
                "algorithmic_tflop_per_launch": kflop / 1e12, "issued_tflop_per_launch": work / 1e12,
                "issued_tflops": issued, "tensor_pipe_frac": issued / peak,
                "traffic_unit": f"bytes/launch (ncu dram read+write); algorithmic bytes = {kbytes:.3g} (16-bit operand planes in and out)",
                "step_tflops": world * FLOP_PER_PACKED_POS * T * T * args.steps / (ms / 1e3) / 1e12}
    del a

    # ---- second view: the HBM-bound kernel class (1x1 conv 128->128 + activation, planes in / planes out: the shape of the
    # ConvTransBlock 1x1 layers, tcm.py:242-268, timed at full resolution).  Algorithmic bytes = operand planes read once + result
    # planes written once.
    roofline_hbm = None
    if args.engine != "fp32":
        try:
            g1 = torch.Generator().manual_seed(3)
            pc1 = ops.pack_weight((torch.randn(128, 128, 1, 1, generator=g1) / 11.3).to(dev), torch.randn(128, generator=g1).to(dev))
            a1 = torch.randn(1, T, T, 128, device=dev)
            sp1 = ops.split_operand(a1, pc1.cp)
            f1 = lambda: ops.conv2d(a1, pc1, act=ops.ACT_LRELU, slope=0.1, presplit=sp1, emit_split=True, keep_fp32=False)
            for _ in range(3):
                f1()
            torch.cuda.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(reps):
                f1()
            h1.record()
            torch.cuda.synchronize()
            hms = h0.elapsed_time(h1) / reps
            per_el = 8 if args.engine == "bf16x3" else 4          # 16-bit hi (+ lo) planes in, hi (+ lo) planes out
            hbytes = float(T) * T * 128 * per_el
            hpeak = float(peaks.get("hbm_gbs", 6650.0))
            roofline_hbm = {"kernel": "conv_tc_kernel -- 1x1 128->128 + LeakyReLU @ full res, operand planes in / out (the per-pixel 1x1 layer class of the path)",
                            "bound": "hbm", "achieved": hbytes / (hms / 1e3) / 1e9, "peak": hpeak, "unit": "GB/s",
                            "frac": hbytes / (hms / 1e3) / 1e9 / hpeak, "traffic": None, "ms_per_launch": hms,
                            "algorithmic_bytes_per_launch": hbytes, "peak_source": peaks["_src"] + ", burst copy bandwidth"}
            del a1, sp1
        except Exception as e:      # an extra view must never cost the bench line
            roofline_hbm = {"error": str(e)[:200]}

    # ---- third view: the fused packed-Bayer ingest kernel of the step (csrc/ingest.cu: lens-shading MLP + conv_first * (lsc + 1),
    # raw2bit.py:1771-1780), timed alone on the step's own inputs.  Algorithmic bytes per pixel: coord 2 x 4 + raw 4 x 4 in, the
    # lens-shading map 128 x 4 (an output of forward()) and the 128-channel bf16 hi + lo operand planes of conv_down out.
    roofline_ingest = None
    if args.engine == "bf16x3":
        try:
            coord_d, raw_d = x_dev[2], ops.to_nhwc(x_dev[0])
            if ops.fused_ingest_ok(model.lsc.layers(), coord_d, model.conv_first):
                f2 = lambda: model.lsc._f_fused(coord_d, raw_d, model.conv_first, emit_stride=2)
                for _ in range(3):
                    f2()
                torch.cuda.synchronize()
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                for _ in range(reps):
                    f2()
                h1.record()
                torch.cuda.synchronize()
                ims = h0.elapsed_time(h1) / reps
                ibytes = float(T) * T * (2 * 4 + 4 * 4 + 128 * 4 + 128 * 2 * 2)
                iflop = float(T) * T * 2 * (2 * 128 + 3 * 128 * 128 + 36 * 128)
                issued_i = float(T) * T * 2 * 3 * (3 * 128 * 128 + 48 * 128)
                hpeak = float(peaks.get("hbm_gbs", 6650.0))
                roofline_ingest = {"kernel": "ingest_kernel<conv> -- lens-shading MLP 2->128->128->128->128 (activations as the tcgen05 A operand "
                                             "in tensor memory) + conv_first 3x3 4->128 * (lsc + 1), one launch",
                                   "bound": "hbm", "achieved": ibytes / (ims / 1e3) / 1e9, "peak": hpeak, "unit": "GB/s",
                                   "frac": ibytes / (ims / 1e3) / 1e9 / hpeak, "ms_per_launch": ims, "algorithmic_bytes_per_launch": ibytes,
                                   "traffic": 0.103809e9 + 4.239047e9 if T == 2048 else None,
                                   "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r2_ingest_ncu.md)",
                                   "algorithmic_tflops": iflop / (ims / 1e3) / 1e12, "issued_tflops": issued_i / (ims / 1e3) / 1e12,
                                   "replaces": "5 launches (4 lens-shading layers + conv_first), 4.89 ms at T=2048",
                                   "peak_source": peaks["_src"] + ", burst copy bandwidth"}
            del raw_d
        except Exception as e:
            roofline_ingest = {"error": str(e)[:200]}

    cpu, mismatch = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Tc = args.cpu_tile or 1024
        r = run_cpu_reference(Tc, 2, 1, 60.0)
        cpu = {"value": r["value"], "unit": "MP/s", "cores": r["threads"], "kind": r["kind"],
               "sample": f"{r['note']}; one 4x{Tc}x{Tc} tile (bounded sample of the 4x{T}x{T} workload), torch CPU fp32, {r['threads']} threads "
                         f"(best of all-cores/32/16 probed on a 4x512x512 tile), 1 warm-up + {r['steps_timed']} timed passes "
                         f"({r['dt']:.2f} s each); host has {os.cpu_count()} logical cores"}
        # parity on the sampled tile, reported with the number: symbols of OUR forward vs the CPU arm's on the same input
        xs = [t.to(dev) for t in inputs.make_inputs(Tc, seed=1234)]
        po = model(xs)["para"]
        ours = torch.round(po["y"] - po["means"]).to(torch.int32).cpu()
        mismatch = {"tile": Tc, "symbols": int(ours.numel()), "mismatching": int((ours != r["sym"]).sum()),
                    "fraction": float((ours != r["sym"]).float().mean()), "vs": r["kind"]}
    if rank == 0:
        dtype = {"fp32": "f32", "bf16x3": "bf16x3 (tcgen05: bf16 hi/lo split operands, 3 MMA passes, fp32 accumulate)", "bf16": "bf16",
                 "fp16": "fp16"}[args.engine]
        if tail:
            dtype += f"; per-stage policy: post-quantisation synthesis tail (raw2bit.py:1680-1682) = {tail} (1 MMA pass, fp32 accumulate)"
        line = {"metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": dtype, "data": "synthetic",
                "config": {"workload": workload_name(T),
                           "tile": T, "tiles_per_gpu": 1, "conv_engine": args.engine, "tail_engine": tail or args.engine,
                           "cuda_graphs": not args.no_graphs, "parallelism": f"tile-sharded x{world}",
                           "l2_policy": "inputs and activations (>2 GB per layer) exceed the 126 MB L2; no explicit flush",
                           "bitstream_bytes": nbytes, "symbols": nsym,
                           "e2e_outputs": "bitstream bytes on the host; x_hat (3 x 2T x 2T fp32) is computed every step but stays on the device",
                           "e2e_inputs": "pinned host -> device copy of every step's inputs, issued one step ahead on a copy stream"},
                "clocks": clocks, "gpu_launches": int(launches),
                "gpu_launches_note": "kernel nodes executed inside the timed region (eager launches + nodes of the replayed CUDA graphs)",
                "e2e": {"value": e2e, "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_ingest": roofline_ingest, "cpu_baseline": cpu,
                "symbols_mismatch_vs_oracle": mismatch, "frame4k": frame4k,
                "decode": {"value": mp_tile / (decode_ms / 1e3), "unit": "MP/s", "ms_per_tile": decode_ms,
                           "what": "decompress(strings, shape) -> x_hat of the same tile (eager launches; host range decoder per slice)"},
                "packed_positions_per_s": value * 0.25e6, "tiles_per_s": value / mp_tile}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
