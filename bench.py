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
def cpu_reference_step(sd, x):
    """The reference's algorithm (oracle restatement) on host cores: forward + range coding of its symbols."""
    import numpy as np
    import torch

    from oracle import refpath

    out = refpath.final_forward(sd, x)
    gc = refpath._gc()
    sym = torch.round(out["para"]["y"] - out["para"]["means"]).to(torch.int32)
    idx = gc.build_indexes(out["para"]["scales"])
    nsl = 5
    N, C, h, w = sym.shape
    # slice-major order, as compress() flattens them (raw2bit.py:1943-1944)
    s = np.concatenate([sym[:, i * C // nsl:(i + 1) * C // nsl].reshape(-1).numpy() for i in range(nsl)])
    i_ = np.concatenate([idx[:, i * C // nsl:(i + 1) * C // nsl].reshape(-1).numpy() for i in range(nsl)])
    return refpath.encode_stream(s, i_, gc)


def run_cpu_reference(T, steps, warmup):
    """Times the oracle port on the host.  torch's CPU kernels do not scale to 100+ threads on these layer sizes,
    so the thread count is probed (all cores, 32, 16) on the warm-up pass and the fastest is used and reported."""
    import torch

    from realcamnet_b200 import synthetic as inputs, synthetic as weights
    from realcamnet_b200 import raw2bit  # parameter names/shapes only; nothing of it runs in this leg

    cores = os.cpu_count() or 1
    m = raw2bit.raw_compression_tcm_final()
    weights.fill_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = inputs.make_inputs(T, seed=1234)
    best, best_t = cores, None
    for th in sorted({cores, min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(th)
        t0 = time.perf_counter()
        cpu_reference_step(sd, x)          # doubles as warm-up
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    torch.set_num_threads(best)
    for _ in range(max(0, warmup - 1)):
        cpu_reference_step(sd, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, x)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    mp = 4.0 * T * T / 1e6
    return mp / dt, dt, best


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tile", type=int, default=2048, help="packed tile side T (BASELINE config[1]: 2048)")
    ap.add_argument("--cpu-tile", type=int, default=1024, help="tile side of the bounded CPU sample (4x1024x1024: ~5 s per pass on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--engine", default=os.environ.get("RCN_CONV_ENGINE", "bf16x3"), choices=["fp32", "bf16x3", "bf16"],
                    help="conv engine: fp32 = CUDA-core exact; bf16x3 = tcgen05 with hi/lo split operands (parity grade); "
                         "bf16 = tcgen05 single pass (fast mode, outside the 1e-3 parity bar)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        Tc = args.cpu_tile
        v, dt, cores = run_cpu_reference(Tc, max(1, args.steps), max(0, min(args.warmup, 1)))
        line = {"metric": METRIC, "value": v, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": f"raw_compression_tcm_final forward+bitstream, bounded sample: one 4x{Tc}x{Tc} tile per step "
                                       f"(ours: 4x{args.tile}x{args.tile})", "tile": Tc},
                "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
                                 "sample": f"oracle restatement of raw2bit.py:1766-1855 + rANS on one 4x{Tc}x{Tc} tile, torch CPU fp32, {cores} threads"},
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

    def step_e2e():
        xs = [t.to(dev, non_blocking=True) for t in x_host]
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
    d2h = 2 * 4 * 320 * (T // 16) ** 2 + 4 * 192 * (T // 64) ** 2     # int32 symbols + indexes (y), symbols (z)

    # ---- roofline of the dominant kernel: the full-resolution 128->128 3x3 conv of the g_s tail (raw2bit.py:1681)
    import ctypes

    from realcamnet_b200 import _C

    peaks = load_peaks()
    pc = ops.pack(model.g_s[10].conv1)
    a = torch.randn(1, T, T, 128, device=dev)
    o = torch.empty_like(a)
    kflop = 2.0 * 9 * 128 * 128 * T * T
    reps = 5
    if args.engine == "fp32":
        fn = lambda: ops.conv2d(a, pc, out=o, act=ops.ACT_LRELU, slope=0.01, engine="fp32")
        kname = "conv2d_kernel<128> (fp32 FFMA implicit GEMM)"
        peak, peak_note = float(peaks.get("bf16_tflops", 1590.0)), "burst bf16 (kernel timed alone); fp32 FFMA peak is ~75 TFLOP/s"
        work = kflop
    else:
        passes = 3 if args.engine == "bf16x3" else 1
        hi = torch.empty(1, T, T, 128, device=dev, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        _C.check(_C.lib().rcn_split_bf16(P(a), 128, T * T, 128, 128, 0, P(hi), P(lo), ops._stream()))
        d = _C.ConvDesc()
        d.x, d.N, d.H, d.W, d.Cin, d.ldx = a.data_ptr(), 1, T, T, 128, 128
        d.w, d.bias, d.k, d.stride, d.Cout = pc.w.data_ptr(), pc.bias.data_ptr(), 3, 1, 128
        d.y, d.ldy, d.store, d.act, d.slope, d.res_scale = o.data_ptr(), 128, 0, ops.ACT_LRELU, 0.01, 1.0
        fn = lambda: _C.check(_C.lib().rcn_conv2d_tc(ctypes.byref(d), P(hi), P(lo), P(pc.w_hi), P(pc.w_lo), 128, passes, ops._stream()))
        kname = f"conv_tc_kernel (tcgen05.mma + TMA, {passes} bf16 pass{'es' if passes > 1 else ''})"
        peak, peak_note = float(peaks.get("bf16_tflops", 1590.0)), "burst bf16 (kernel timed alone)"
        work = kflop * passes          # tensor-pipe FLOPs actually issued (hi*hi + lo*hi + hi*lo for bf16x3)
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
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the one `ncu --set full` capture of THIS launch at
    # T=2048 (profiles/r1_final_conv_tcgen05_ncu_full.md: 2.150 + 2.109 GB; profiles/r1_conv_fp32_ncu_full.md: 2.155 + 2.102 GB);
    # other tile sizes were not captured.
    traffic = {("bf16x3", 2048): 4.259602e9, ("fp32", 2048): 4.257296e9}.get((args.engine, T))
    roofline = {"kernel": kname + " -- 3x3 128->128 @ full res (g_s tail)", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "ms_per_launch": kms, "peak_source": peaks["_src"] + ", " + peak_note,
                "algorithmic_tflop_per_launch": kflop / 1e12, "issued_tflop_per_launch": work / 1e12,
                "issued_tflops": issued, "tensor_pipe_frac": issued / peak,
                "traffic_unit": "bytes/launch (ncu dram read+write); algorithmic bytes = 4.29e9 (bf16 hi+lo planes in, fp32 map out)",
                "step_tflops": world * FLOP_PER_PACKED_POS * T * T * args.steps / (ms / 1e3) / 1e12}
    del a, o

    # ---- second view: the dominant HBM-bound kernel class (1x1 conv 128->128 + LeakyReLU at full resolution, planes in / planes out:
    # the lens-shading MLP layers, LiteISP.py:363-378).  Algorithmic bytes = operand planes read once + result planes written once.
    roofline_hbm = None
    if args.engine != "fp32":
        try:
            g1 = torch.Generator().manual_seed(3)
            pc1 = ops.pack_weight((torch.randn(128, 128, 1, 1, generator=g1) / 11.3).to(dev), torch.randn(128, generator=g1).to(dev))
            a1 = torch.randn(1, T, T, 128, device=dev)
            sp1 = ops.split_operand(a1, pc1.cp, passes=3 if args.engine == "bf16x3" else 1)
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
            per_el = 8 if args.engine == "bf16x3" else 4          # bf16 hi (+ lo) in, bf16 hi (+ lo) out
            hbytes = float(T) * T * 128 * per_el
            hpeak = float(peaks.get("hbm_gbs", 6650.0))
            roofline_hbm = {"kernel": "conv_tc_kernel -- 1x1 128->128 + LeakyReLU @ full res, operand planes in / out (lens-shading MLP layer)",
                            "bound": "hbm", "achieved": hbytes / (hms / 1e3) / 1e9, "peak": hpeak, "unit": "GB/s",
                            "frac": hbytes / (hms / 1e3) / 1e9 / hpeak, "traffic": None, "ms_per_launch": hms,
                            "algorithmic_bytes_per_launch": hbytes, "peak_source": peaks["_src"] + ", burst copy bandwidth"}
            del a1, sp1
        except Exception as e:      # an extra view must never cost the bench line
            roofline_hbm = {"error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = run_cpu_reference(args.cpu_tile, 2, 1)
        cpu = {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": f"oracle restatement (raw2bit.py:1766-1855 + rANS) on one 4x{args.cpu_tile}x{args.cpu_tile} tile, "
                         f"torch CPU fp32, {cores} threads (best of all-cores/32/16 probed on the warm-up), 2 timed passes ({dt:.2f} s each); host has {os.cpu_count()} logical cores"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (tcgen05, fp32 accumulate)", "bf16": "bf16"}[args.engine], "data": "synthetic",
                "config": {"workload": f"raw_compression_tcm_final.forward + range coder on one 4x{T}x{T} packed-Bayer tile per GPU "
                                       "(BASELINE config[1]), random-init weights (name-keyed, seed 0)",
                           "tile": T, "tiles_per_gpu": 1, "conv_engine": args.engine, "cuda_graphs": not args.no_graphs, "parallelism": f"tile-sharded x{world}",
                           "l2_policy": "inputs and activations (>2 GB per layer) exceed the 126 MB L2; no explicit flush",
                           "bitstream_bytes": nbytes, "symbols": nsym},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
                "packed_positions_per_s": value * 0.25e6, "tiles_per_s": value / mp_tile}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
