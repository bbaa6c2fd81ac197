#!/usr/bin/env python
"""bench.py — BN254 KZG commit + coset LDE throughput (BASELINE.json metric) on N B200s.

One "step" = the hot path over ONE synthetic trace of 2^20 rows x 16 columns (BASELINE.json configs[1]):
    KzgPcs::commit                 = coset iDFT (2^20 x 16) + 16 G1 MSMs of 2^20 points   (kzg/src/pcs.rs:223-265)
    get_evaluations_on_domain      = zero-pad + coset NTT onto the 2^21-point coset 5*K      (blow-up 2)
value = cols*rows / s with the trace resident in HBM (CUDA events on the launch stream, max over ranks).
e2e   = the same through the host-buffer C ABI: pinned host trace -> H2D -> commit -> LDE -> D2H.
N > 1 ("strong"): the SAME 2^20 x 16 trace, columns sharded over the ranks (rank r: column_shard(16, N, r); every
column's iDFT / LDE / MSM is independent, kzg/src/pcs.rs:244-249); the only exchange is an all_gather of the
commitments (64 B per column).  Side objects of the same line: `weak` (every rank its own 16 columns), `open`
(KzgPcs::open at 2 points), `msm_2p24` (standalone 2^24-point MSM, point-index sharded, BASELINE configs[2]),
`mctx` (rank 0 alone drives all N GPUs through ONE eon_mctx: what a single-process Rust prover would call),
`parity_ok` (full-size identities checked before timing: commit == p(alpha) G, LDE rows == Horner values,
sharded == whole).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl eon|reference] [--log-rows 20] [--cols 16]

--impl reference times the CPU port of the reference path (oracle/c, OpenMP, all host threads) on the same
workload; the reference itself is Rust + halo2curves and cannot be built here.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kzg_commit_lde_cols_rows_per_s"
UNIT = "cols*rows/s"
ALPHA = 12345          # kzg-example/examples/fibonacci_kzg.rs:79
SHIFT_LDE = 5          # Fr::GENERATOR: quotient domain 5*K (commit/src/domain.rs:167)
IMAD_PER_MODMUL = 272  # 136 32-bit limb MACs, lo + hi
MODMUL_PER_MIXED_ADD = 10
DTYPE = "u256 (8x32-bit Montgomery limbs, BN254 Fr/Fq)"
P_TOP = 0x30644e72e131a029  # top 64-bit limb of the Fr modulus
MSM_WINDOW_BITS = -1        # msm workload: forced window bits of the (shard) tables, -1 = the library's cost model


def shared_config(args):
    """The workload, in the same words for both arms (the driver compares the two `config` objects)."""
    return {"workload": f"KZG commit (coset iDFT + {args.cols} MSM over a 2^{args.log_rows}-point SRS) + "
                        f"blow-up-{1 << args.added_bits} coset LDE of one 2^{args.log_rows} x {args.cols} BN254 Fr trace "
                        "(BASELINE configs[1] shape family)",
            "rows": 1 << args.log_rows, "cols": args.cols, "added_bits": args.added_bits,
            "srs_points": 1 << args.log_rows, "lde_shift": SHIFT_LDE, "seed": "column c = rng(1000 + c)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """DRAM bytes per step of a kernel family, from this round's ncu capture (profiles/traffic.json, written by
    tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(key)
        return (float(e["bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Sampled in-process through
    NVML on a thread (a `nvidia-smi -lms` child needs longer to start than 5 steps of 43 ms take, and would come
    back with no samples); the nvidia-smi loop remains as the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, dev, period_s=0.01):
        self.dev = dev
        self.period = period_s
        self.proc = None
        self.thread = None
        self.sm, self.power, self.reasons = [], [], set()
        self.mx = None
        self.nv = None
        self.handle = None
        self.source = None
        try:  # set up before the timed region: nvmlInit takes tens of ms
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index(pynvml, dev))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.masks = {
                "hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap,
            }
        except Exception:
            self.nv = None

    @staticmethod
    def _nvml_index(pynvml, dev):
        """NVML enumerates every GPU of the box; CUDA only the visible ones, in CUDA_VISIBLE_DEVICES order."""
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if dev < len(ids) and ids[dev].isdigit():
                return int(ids[dev])
        return dev

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
        except Exception:
            pass
        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for n, m in self.masks.items():
            if r & m:
                self.reasons.add(n)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            import threading
            self._stop = threading.Event()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            self.source = "nvml thread, %d ms period" % int(self.period * 1000)
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            try:
                self._sample()      # at least one sample taken with the last kernels still in flight / just retired
            except Exception:
                pass
            self._stop.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "sm_min_mhz": min(self.sm) if self.sm else None,
                    "power_w_max": max(self.power) if self.power else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": self.source}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": self.source}


def synth_column(c, rows):
    """Column c of the synthetic trace: uniform Fr in Montgomery form exactly like the reference sampler
    (bn254/src/field.rs:534-551): random 256 bits, top 2 bits cleared, rejected if >= P, used AS the Montgomery
    limbs.  Seeded per column, so a rank can build just its shard."""
    from plonky3_eon_b200 import field
    rng = np.random.default_rng(1000 + c)
    out = rng.integers(0, 1 << 64, size=(rows, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 62) - 1)
    pl = [(field.P >> (64 * i)) & ((1 << 64) - 1) for i in range(4)]
    while True:
        lt = np.zeros(rows, dtype=bool)
        eq = np.ones(rows, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (out[:, k] < np.uint64(pl[k]))
            eq &= out[:, k] == np.uint64(pl[k])
        bad = np.nonzero(~lt)[0]
        if len(bad) == 0:
            break
        rep = rng.integers(0, 1 << 64, size=(len(bad), 4), dtype=np.uint64)
        rep[:, 3] &= np.uint64((1 << 62) - 1)
        out[bad] = rep
    return out


def synth_trace(rows, cols, c0=0):
    out = np.empty((rows, cols, 4), dtype=np.uint64)
    for j in range(cols):
        out[:, j] = synth_column(c0 + j, rows)
    return out


def synth_device(rows, cols, c0, dev, seed=1000):
    """The same distribution generated on the device (large shapes: 2^24 rows): top limb < the modulus's top limb
    (the equality case, probability 2^-62, is rejected as well)."""
    import torch
    out = torch.empty((rows, cols, 4), dtype=torch.int64, device=dev)
    for j in range(cols):
        g = torch.Generator(device=dev)
        g.manual_seed(seed + c0 + j)
        x = torch.randint(-(1 << 63), (1 << 63) - 1, (rows, 4), generator=g, dtype=torch.int64, device=dev)
        top = x[:, 3] & ((1 << 62) - 1)
        bad = top >= P_TOP
        while bool(bad.any()):
            k = int(bad.sum())
            rep = torch.randint(0, (1 << 62) - 1, (k,), generator=g, dtype=torch.int64, device=dev)
            top[bad] = rep
            bad = top >= P_TOP
        x[:, 3] = top
        out[:, j] = x
    return out


def column_shard(width, world, rank):
    base, extra = divmod(width, world)
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPU cores next to its GPU before any pinned host buffer is allocated
    (first touch then places the buffers on that NUMA node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(log_rows, cols, added_bits, srs, ev, msm_cols=None):
    """One step of the CPU port (oracle/c): coset iDFT of all columns, MSM of `msm_cols` columns (all by
    default), zero-pad + coset DFT.  Returns (seconds with the MSM part scaled to all columns, detail)."""
    from oracle import cport
    rows = 1 << log_rows
    msm_cols = cols if msm_cols is None else msm_cols
    t0 = time.perf_counter()
    _, coeffs = cport.kzg_commit(ev, 1, srs, ncols_msm=0)                      # coset iDFT only
    t1 = time.perf_counter()
    cport.msm(srs, coeffs, ncols=msm_cols, ld=cols)
    t2 = time.perf_counter()
    pad = np.zeros((rows << added_bits, cols, 4), dtype=np.uint64)
    pad[:rows] = coeffs
    t3 = time.perf_counter()
    cport.coset_dft_batch(pad, SHIFT_LDE)                                      # zero-pad + coset DFT
    t4 = time.perf_counter()
    est = (t1 - t0) + (t2 - t1) * (cols / msm_cols) + (t4 - t3)
    return est, {"idft_s": t1 - t0, "msm_s": t2 - t1, "msm_cols_run": msm_cols, "lde_s": t4 - t3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone and is meant to use
    # every host core it can get (the reference's rayon pool would), so undo that before OpenMP starts
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    from oracle import cport
    cores = cport.num_threads()
    rows = 1 << args.log_rows
    t_setup = time.perf_counter()
    srs = cport.srs_generate(ALPHA, rows)
    ev = synth_trace(rows, args.cols)
    t_setup = time.perf_counter() - t_setup
    steps = max(1, args.steps)
    warm = args.warmup_ref if args.warmup_ref > 0 else max(0, args.warmup)
    runs = []
    for i in range(warm + steps):
        est, d = cpu_reference_step(args.log_rows, args.cols, args.added_bits, srs, ev)   # every column's MSM is run
        if i >= warm:
            runs.append((est, d))
    est = float(np.median([x[0] for x in runs]))
    v = rows * args.cols / est
    sample = (f"the whole step, nothing extrapolated: coset iDFT 2^{args.log_rows}x{args.cols}, MSM of all {args.cols} "
              f"columns, coset LDE to 2^{args.log_rows + args.added_bits} rows; affine SRS normalised once (the "
              "reference re-normalises per call, bn254/src/curve.rs:170: this favours the baseline)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery, BN254 Fr/Fq)", "data": "synthetic",
        "config": shared_config(args),
        "config_detail": {"arm": "CPU port of the reference path (oracle/c, OpenMP); the Rust reference cannot be built "
                                 "in this image", "srs_and_trace_setup_s": t_setup},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "detail": runs[-1][1]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ------------------------------------------------------------------------------------------------
class Harness:
    """torch.distributed / CUDA plumbing shared by the workloads."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the eon arm has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.cpus = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        self.host_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            # a second, CPU-side group: ranks that only WAIT (while rank 0 drives every GPU through one eon_mctx) must
            # not sit in an NCCL barrier, whose spinning kernel would time-slice their GPU with rank 0's work
            self.host_group = dist.new_group(backend="gloo")
        self.stream = torch.cuda.current_stream()

    def host_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.host_group)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """`steps` calls bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def all_ok(self, ok):
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def gather_u64(self, arr):
        """all_gather of a uint64 numpy array (same shape everywhere) through the device -> list per rank."""
        torch = self.torch
        if self.world == 1:
            return [arr]
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).reshape(-1).copy()).to(self.dev)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return [o.cpu().numpy().view(np.uint64).reshape(arr.shape) for o in outs]

    def done(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def fr_to_int(w):
    from plonky3_eon_b200 import field
    return field.from_wire(w)


def check_commit_identity(ctx, d_coeffs_ptr, rows, cols, ld, commits, C):
    """commit_c == p_c(alpha) * G for the synthetic SRS g1_powers[i] = alpha^i G (kzg/src/params.rs:123-139):
    p_c(alpha) by the library's Horner scan (csrc/kzg.cu), times G by a 1-point MSM over SRS[0] = G — kernels
    disjoint from the ones that produced `commits` (coset iDFT is shared with neither side of the check: the
    coefficients come from the commit itself)."""
    import torch
    from plonky3_eon_b200 import field
    assert ld == cols
    d_quot = torch.empty((rows, cols, 4), dtype=torch.int64, device="cuda")
    vals = np.zeros((cols, 4), dtype=np.uint64)
    ctx.call("eon_quotient_and_eval_dev", C.c_void_p(d_coeffs_ptr), rows, cols, field.to_wire(ALPHA),
             C.c_void_p(d_quot.data_ptr()), vals)
    del d_quot
    want = np.zeros((cols, 8), dtype=np.uint64)
    ctx.call("eon_msm_srs", vals, 1, cols, cols, want)
    return bool(np.array_equal(want, commits)), vals


def check_lde_rows(ctx, d_coeffs_ptr, rows, cols, d_lde, log_lde, C, sample_rows):
    """LDE row j == (p_c(5 * omega^j))_c, the Horner value of the reference's own get_evaluations_on_domain
    (kzg/src/pcs.rs:278-286), by the library's scan kernel."""
    import torch
    from plonky3_eon_b200 import field
    d_quot = torch.empty((rows, cols, 4), dtype=torch.int64, device="cuda")
    w = field.two_adic_generator(log_lde)
    ok = True
    for j in sample_rows:
        x = SHIFT_LDE * pow(w, j, field.P) % field.P
        vals = np.zeros((cols, 4), dtype=np.uint64)
        ctx.call("eon_quotient_and_eval_dev", C.c_void_p(d_coeffs_ptr), rows, cols, field.to_wire(x),
                 C.c_void_p(d_quot.data_ptr()), vals)
        got = d_lde[j].cpu().numpy().view(np.uint64)
        ok = ok and bool(np.array_equal(got, vals))
    del d_quot
    return ok


def imad_roofline(ctx, phases, steps, adds, rounds, label, traffic_key):
    """IMAD roofline of the dominant kernel family of an MSM-bound step: the k_tree_bwd launches (batched-affine
    pair additions, 5 Fq products each), or the XYZZ accumulation when no rounds ran."""
    imad_peak = max(ctx.imad_peak_tops(0), ctx.imad_peak_tops(1))       # T IMAD/s, measured now
    hbm_peak, _ = peaks()
    common = {
        "bound": "imad", "peak": imad_peak, "unit": "TIMAD/s",
        "peak_source": "eon_bench_imad_peak in this run (mad.lo/mad.hi.u32, 16 independent chains/thread); IMAD is not "
                       "in MEASURED_PEAKS.json",
        "note": "integer-pipe bound (SURVEY §8d): the roofline is IMAD, not HBM/tensor",
    }
    acc_ms = phases["msm_accumulate"] / steps
    traffic, tsrc = ncu_traffic(traffic_key)
    if rounds:
        pairs = sum(adds // (1 << (r + 1)) for r in range(rounds))
        bwd_ms = phases["msm_tree_bwd"] / steps
        ops = pairs * 5 * IMAD_PER_MODMUL
        achieved = ops / (bwd_ms * 1e-3) / 1e12
        alg_bytes = pairs * 224    # per pair: 2 points in (128 B), prefix product in (32 B), sum out (64 B)
        return dict(common, kernel=f"k_tree_bwd x{rounds} ({label}: batched-affine pair additions, 5 modmul per pair)",
                    achieved=achieved, frac=achieved / imad_peak, algorithmic_ops_per_launch=ops, launch_ms=bwd_ms,
                    traffic=traffic, traffic_source=tsrc,
                    hbm_view={"bound": "hbm", "algorithmic_bytes": alg_bytes, "peak": hbm_peak, "unit": "GB/s",
                              "achieved": alg_bytes / (bwd_ms * 1e-3) / 1e9,
                              "frac": alg_bytes / (bwd_ms * 1e-3) / 1e9 / hbm_peak},
                    accumulate_phase={"ms": acc_ms, "tree_fwd_ms": phases["msm_tree_fwd"] / steps,
                                      "tree_inv_ms": phases["msm_tree_inv"] / steps, "tree_bwd_ms": bwd_ms,
                                      "finish_ms": phases["msm_finish"] / steps, "bucket_additions": adds,
                                      "xyzz_equiv_frac": adds * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL
                                      / (acc_ms * 1e-3) / 1e12 / imad_peak})
    ops = adds * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL
    achieved = ops / (acc_ms * 1e-3) / 1e12
    return dict(common, kernel=f"k_msm_accumulate ({label}: XYZZ mixed adds, one thread per bucket)", achieved=achieved,
                frac=achieved / imad_peak, algorithmic_ops_per_launch=ops, launch_ms=acc_ms, traffic=traffic,
                traffic_source=tsrc)


# ------------------------------------------------------------------------------------------------
def msm_leg(H, ctx, log_n, cols, steps, warmup, C):
    """BASELINE configs[2] at one size: 2^log_n points x `cols` scalar columns of random Fr over the synthetic SRS,
    points sharded by index range over the ranks ("strong").  Every rank keeps window tables for ITS range only
    (sized for the shard), leaves its partial sums on the device, NCCL all_gathers them (world x cols x 64 B) and
    one launch adds them — no host hop."""
    import torch
    from plonky3_eon_b200 import field
    n = 1 << log_n
    world, rank = H.world, H.rank
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    cnt = base + (1 if rank < extra else 0)
    ctx.call("eon_srs_generate_unsafe", field.to_wire(ALPHA), n)
    if world > 1:
        ctx.call("eon_srs_set_window_tables", 0)            # the whole-SRS tables are not needed: shard tables below
        ctx.call("eon_srs_set_range_tables", first, cnt, max(MSM_WINDOW_BITS, 0))
    elif MSM_WINDOW_BITS >= 0:
        ctx.call("eon_srs_set_window_tables", MSM_WINDOW_BITS)
    d_sc = synth_device(cnt, cols, 0, H.dev, seed=4242 + 7919 * rank)   # seed 42: bn254/benches/bench_curve.rs:40
    d_part = torch.zeros((cols, 8), dtype=torch.int64, device=H.dev)
    d_all = torch.zeros((world, cols, 8), dtype=torch.int64, device=H.dev)
    out = np.zeros((cols, 8), dtype=np.uint64)

    def step():
        if world == 1:
            ctx.call("eon_msm_srs_dev", C.c_void_p(d_sc.data_ptr()), n, cols, cols, out)
            return
        ctx.call("eon_msm_srs_range_partial_dev", C.c_void_p(d_sc.data_ptr()), first, cnt, cols, cols,
                 C.c_void_p(d_part.data_ptr()))
        H.dist.all_gather_into_tensor(d_all, d_part)                      # same stream as the context: ordered
        ctx.call("eon_g1_sum_cols_dev", C.c_void_p(d_all.data_ptr()), world, cols, out)

    step()
    # parity at full size: sum_i s_i alpha^i by the Horner scan of every shard, combined on the host, times G
    d_quot = torch.empty((cnt, cols, 4), dtype=torch.int64, device=H.dev)
    vals = np.zeros((cols, 4), dtype=np.uint64)
    ctx.call("eon_quotient_and_eval_dev", C.c_void_p(d_sc.data_ptr()), cnt, cols, field.to_wire(ALPHA),
             C.c_void_p(d_quot.data_ptr()), vals)
    del d_quot
    parts = H.gather_u64(vals)
    total = [0] * cols
    for r in range(world):
        f_r = r * base + min(r, extra)
        scale = pow(ALPHA, f_r, field.P)
        for c in range(cols):
            total[c] = (total[c] + fr_to_int(parts[r][c]) * scale) % field.P
    want = np.zeros((cols, 8), dtype=np.uint64)
    ctx.call("eon_msm_srs", np.stack([field.to_wire(t) for t in total]), 1, cols, cols, want)
    parity = H.all_ok(bool(np.array_equal(want, out)))

    for _ in range(warmup):
        step()
    ctx.phase_reset()
    launches0 = ctx.launch_count()
    ms = H.timed(step, steps)
    launches = ctx.launch_count() - launches0
    phases = ctx.phase_ms()
    res = {"metric": "msm_points_per_s", "value": n * cols * steps / (ms * 1e-3), "unit": "points/s",
           "ms_per_step": ms / steps, "points": n, "cols": cols, "n_gpus": world, "scaling": "strong",
           "parity_ok": parity, "gpu_launches": launches,
           "parity": "result == (sum_i s_i alpha^i) G, the sum by the Horner-scan kernel per shard",
           "sharding": "one GPU" if world == 1 else f"point index ranges over {world} ranks, per-shard window tables, "
                       "ncclAllGather of the partial sums + one add launch",
           "phase_ms_per_step": {k: v / steps for k, v in phases.items()}}
    res["msm_affine_rounds"] = int(ctx.lib.eon_msm_rounds_used(ctx.h))
    res["msm_window_bits"] = int(ctx.lib.eon_msm_window_bits_used(ctx.h))   # of the shard's tables
    return res, phases


def open_leg(H, ctx, handle, log_rows, cols, steps, warmup, C):
    """KzgPcs::open (kzg/src/pcs.rs:289-335) of this rank's columns at the two points the prover uses (zeta,
    zeta * omega; eon-uni-stark/src/prover.rs:416-431): 2 * cols synthetic divisions and ONE batched MSM of
    2 * cols columns; points in, values and witnesses out through host buffers (so device time == e2e here)."""
    from plonky3_eon_b200 import field
    rows = 1 << log_rows
    zeta = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    omega = field.two_adic_generator(log_rows)
    pts = np.stack([field.to_wire(zeta), field.to_wire(zeta * omega % field.P)])
    vals = np.zeros((2, cols, 4), dtype=np.uint64)
    wits = np.zeros((2, cols, 8), dtype=np.uint64)

    def step():
        ctx.call("eon_kzg_open", handle, pts, 2, vals, wits)
        if H.world > 1:
            H.gather_u64(wits)

    for _ in range(max(1, warmup)):
        step()
    ctx.phase_reset()
    launches0 = ctx.launch_count()
    ms = H.timed(step, steps)
    launches = ctx.launch_count() - launches0
    phases = ctx.phase_ms()
    return {"ms": ms, "launches": launches, "phases": phases, "vals": vals.copy(), "wits": wits.copy(), "pts": pts,
            "zeta": zeta, "omega": omega, "rows": rows, "rounds": int(ctx.lib.eon_msm_rounds_used(ctx.h))}


# ------------------------------------------------------------------------------------------------
def run_eon(args):
    import ctypes as C

    import torch

    import plonky3_eon_b200 as eon
    from plonky3_eon_b200 import field

    H = Harness()
    world, rank, local = H.world, H.rank, H.local
    rows, log_rows, cols_total, ab = 1 << args.log_rows, args.log_rows, args.cols, args.added_bits
    c0, c1 = column_shard(cols_total, world, rank)
    cols = c1 - c0
    if cols == 0:
        raise SystemExit(f"bench.py: {cols_total} columns cannot be sharded over {world} ranks")
    maxw = column_shard(cols_total, world, 0)[1]

    ctx = eon.Context(local, stream=H.stream.cuda_stream)
    pcs = eon.GpuKzgPcs.new(rows - 1, ALPHA, ctx=ctx)           # synthetic SRS alpha^i * G on the device
    if args.window_bits >= 0:                                    # default: the library's own table policy
        ctx.call("eon_srs_set_window_tables", args.window_bits)
    if args.slice_schedule >= 0:
        ctx.call("eon_msm_set_slice_schedule", args.slice_schedule)
    if args.msm_rounds >= 0:
        ctx.call("eon_msm_set_rounds", args.msm_rounds)
    shift_one = field.to_wire(1)
    shift_lde = field.to_wire(SHIFT_LDE)

    # ---- the synthetic trace: whole matrix pinned on the host (every rank, e2e reads its columns out of it),
    # this rank's columns resident on the device ---------------------------------------------------------------
    if args.no_e2e:
        host_pin_np = lde_pin_np = None
        d_evals = synth_device(rows, cols, c0, H.dev)
    else:
        host_pin = torch.empty((rows, cols_total, 4), dtype=torch.int64).pin_memory()
        host_pin_np = host_pin.numpy().view(np.uint64)
        host_pin_np[:] = synth_trace(rows, cols_total)
        lde_pin = torch.empty((rows << ab, cols_total, 4), dtype=torch.int64).pin_memory()
        lde_pin_np = lde_pin.numpy().view(np.uint64)
        d_evals = host_pin[:, c0:c1].contiguous().to(H.dev)
    d_lde = torch.empty((rows << ab, cols, 4), dtype=torch.int64, device=H.dev)
    commits = np.zeros((cols, 8), dtype=np.uint64)
    pad = np.zeros((maxw, 8), dtype=np.uint64)
    d_pad = torch.zeros(maxw * 8, dtype=torch.int64, device=H.dev)
    gathered = torch.zeros(world * maxw * 8, dtype=torch.int64, device=H.dev)

    def gather_commits():
        if world > 1:                                            # 64 B per column, the only exchange of the step
            pad[:cols] = commits
            d_pad.copy_(torch.from_numpy(pad.view(np.int64).reshape(-1)))
            H.dist.all_gather_into_tensor(gathered, d_pad)

    def step_device(free=True):
        # commit + the hinted quotient-coset LDE in one call (the LDE transform runs beside the MSM)
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_lde_dev", C.c_void_p(d_evals.data_ptr()), log_rows, cols, shift_one, commits,
                 C.byref(h), log_rows + ab, shift_lde, C.c_void_p(d_lde.data_ptr()))
        if free:
            ctx.call("eon_handle_free", h)
        gather_commits()
        return h

    def step_device_two_calls():
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_dev", C.c_void_p(d_evals.data_ptr()), log_rows, cols, shift_one, commits,
                 C.byref(h))
        ctx.call("eon_kzg_evals_on_coset_dev", h, log_rows + ab, shift_lde, C.c_void_p(d_lde.data_ptr()))
        ctx.call("eon_handle_free", h)
        gather_commits()

    def step_e2e():
        # what a Pcs shim with an LDE hint calls from commit(): this rank's columns straight out of the caller's
        # row-major host matrix (row pitch = all columns) and its LDE columns straight back into the host result
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_lde_ld", C.c_void_p(host_pin_np.ctypes.data + c0 * 32), cols_total, log_rows, cols,
                 shift_one, commits, C.byref(h), log_rows + ab, shift_lde,
                 C.c_void_p(lde_pin_np.ctypes.data + c0 * 32), cols_total)
        ctx.call("eon_handle_free", h)
        gather_commits()

    def step_e2e_two_calls():
        # the unhinted trait sequence: Pcs::commit, then Pcs::get_evaluations_on_domain
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_ld", C.c_void_p(host_pin_np.ctypes.data + c0 * 32), cols_total, log_rows, cols,
                 shift_one, commits, C.byref(h))
        ctx.call("eon_kzg_evals_on_coset_ld", h, log_rows + ab, shift_lde, C.c_void_p(lde_pin_np.ctypes.data + c0 * 32),
                 cols_total)
        ctx.call("eon_handle_free", h)
        gather_commits()

    # ---- parity at full size, before anything is timed --------------------------------------------------------
    parity = {}
    h0 = step_device(free=False)
    d_coeffs = torch.empty((rows, cols, 4), dtype=torch.int64, device=H.dev)
    ctx.call("eon_coset_idft_batch_dev", C.c_void_p(d_evals.data_ptr()), C.c_void_p(d_coeffs.data_ptr()), log_rows, cols,
             shift_one)
    ok_commit, _ = check_commit_identity(ctx, d_coeffs.data_ptr(), rows, cols, cols, commits, C)
    n_lde = rows << ab
    ok_lde = check_lde_rows(ctx, d_coeffs.data_ptr(), rows, cols, d_lde, log_rows + ab, C,
                            [0, 1, n_lde // 2 + 3, n_lde - 1])
    del d_coeffs
    parity["commit_eq_p_alpha_G"] = H.all_ok(ok_commit)
    parity["lde_rows_eq_horner"] = H.all_ok(ok_lde)
    commits_first = commits.copy()
    if world > 1 and not args.no_e2e:
        # sharded == whole: rank 0 commits all columns on its own GPU and compares with the gathered shards
        H.barrier()
        got = gathered.cpu().numpy().view(np.uint64).reshape(world, maxw, 8)
        ok = True
        if rank == 0:
            d_all = host_pin.to(H.dev)
            whole = np.zeros((cols_total, 8), dtype=np.uint64)
            hw = C.c_uint64(0)
            ctx.call("eon_kzg_commit_dev", C.c_void_p(d_all.data_ptr()), log_rows, cols_total, shift_one, whole, C.byref(hw))
            ctx.call("eon_handle_free", hw)
            del d_all
            for r in range(world):
                a, b = column_shard(cols_total, world, r)
                ok = ok and bool(np.array_equal(got[r][:b - a], whole[a:b]))
        parity["sharded_eq_whole"] = H.all_ok(ok)

    # ---- device-resident legs ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    ctx.phase_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ms_dev = H.timed(step_device, args.steps)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    phases_fused = ctx.phase_ms()          # summed over the timed steps
    parity["steps_repeatable"] = H.all_ok(bool(np.array_equal(commits, commits_first)))
    lde_dev_fused = d_lde.clone() if args.check_e2e else None
    step_device_two_calls()
    ctx.phase_reset()
    ms_dev2 = H.timed(step_device_two_calls, args.steps)
    phases = ctx.phase_ms()   # the same kernels back to back on one stream: per-kernel times for the rooflines
    parity["two_calls_eq_fused"] = H.all_ok(bool(np.array_equal(commits, commits_first))
                                            and (lde_dev_fused is None or bool(torch.equal(lde_dev_fused, d_lde))))
    del lde_dev_fused
    c_bits = int(ctx.lib.eon_msm_window_bits_used(ctx.h))
    W = (255 + c_bits - 1) // c_bits
    rounds_commit = int(ctx.lib.eon_msm_rounds_used(ctx.h))

    # ---- host-buffer legs -------------------------------------------------------------------------------------------
    ms_rowblock = None
    if args.no_e2e:
        ms_e2e = ms_e2e2 = None
    else:
        for _ in range(max(1, args.warmup)):                    # the same W untimed steps as the device leg
            step_e2e()
        ctx.phase_reset()
        ms_e2e = H.timed(step_e2e, args.steps)
        phases_e2e = ctx.phase_ms()    # GPU time of every phase inside the host-buffer steps (streams overlap)
        ok = bool(np.array_equal(commits, commits_first))
        # the LDE that came back over PCIe == the device-resident one (this rank's columns of the host result)
        ok = ok and bool(np.array_equal(lde_pin_np[:, c0:c1], d_lde.cpu().numpy().view(np.uint64)))
        step_e2e_two_calls()
        ms_e2e2 = H.timed(step_e2e_two_calls, args.steps)
        ok = ok and bool(np.array_equal(commits, commits_first))
        parity["e2e_eq_device"] = H.all_ok(ok)
        # few columns per rank: contiguous row blocks over PCIe, the column <-> row exchange over NVLink
        if world > 1 and cols <= args.rowblock_max_cols and rows % world == 0 and cols_total % world == 0:
            from plonky3_eon_b200 import dist as edist
            rb_pipe = edist.RowBlockCommitLde(ctx, log_rows, cols_total, ab, H.dev)
            lde_pin.zero_()

            def step_e2e_rowblock():
                h = rb_pipe.step(host_pin, lde_pin, shift_one, shift_lde, commits)
                ctx.call("eon_handle_free", h)
                gather_commits()
            for _ in range(max(1, args.warmup)):
                step_e2e_rowblock()
            ms_rowblock = H.timed(step_e2e_rowblock, args.steps)
            torch.cuda.synchronize()
            ok = bool(np.array_equal(commits, commits_first))
            sums = rb_pipe.shard_checksums()                                   # [dest rank] of my columns
            all_sums = torch.empty(world * world, dtype=torch.int64, device=H.dev)
            H.dist.all_gather_into_tensor(all_sums, sums)
            all_sums = all_sums.view(world, world).cpu().numpy()              # [src rank][dest rank]
            lrb = (rows << ab) // world
            mine = lde_pin.numpy()[rank * lrb:(rank + 1) * lrb]                 # my row block, all columns
            for j in range(world):
                a, b_ = column_shard(cols_total, world, j)
                with np.errstate(over="ignore"):
                    got = np.int64(mine[:, a:b_].sum(dtype=np.int64))
                ok = ok and bool(got == all_sums[j][rank])
            ok = ok and bool(np.array_equal(mine[:, c0:c1].view(np.uint64),
                                            d_lde[rank * lrb:(rank + 1) * lrb].cpu().numpy().view(np.uint64)))
            parity["rowblock_e2e_eq_device"] = H.all_ok(ok)
            del rb_pipe

    units = rows * cols_total * args.steps
    value = units / (ms_dev * 1e-3)
    ms_e2e_strided = ms_e2e
    if ms_rowblock is not None and ms_rowblock < ms_e2e:
        ms_e2e = ms_rowblock           # the faster transport is what a caller would use at this shard width
    e2e_value = units / (ms_e2e * 1e-3) if ms_e2e else None

    # ---- open: KzgPcs::open of the committed trace at (zeta, zeta * omega) ----------------------------------------
    open_obj = None
    if not args.no_open:
        o = open_leg(H, ctx, h0, log_rows, cols, args.steps, args.warmup, C)
        # parity: the opened values are the Horner values of the same polynomials (checked through the LDE-style
        # identity value == p(z)) and every witness satisfies W == (p(alpha) - p(z)) / (alpha - z) * G
        d_coeffs = torch.empty((rows, cols, 4), dtype=torch.int64, device=H.dev)
        ctx.call("eon_coset_idft_batch_dev", C.c_void_p(d_evals.data_ptr()), C.c_void_p(d_coeffs.data_ptr()), log_rows,
                 cols, shift_one)
        _, p_alpha = check_commit_identity(ctx, d_coeffs.data_ptr(), rows, cols, cols, commits_first, C)
        del d_coeffs
        ok = True
        want_sc = np.zeros((2, cols, 4), dtype=np.uint64)
        for p, z in enumerate((o["zeta"], o["zeta"] * o["omega"] % field.P)):
            inv = pow((ALPHA - z) % field.P, -1, field.P)
            for c in range(cols):
                q = (fr_to_int(p_alpha[c]) - fr_to_int(o["vals"][p, c])) * inv % field.P
                want_sc[p, c] = field.to_wire(q)
        want = np.zeros((2 * cols, 8), dtype=np.uint64)
        ctx.call("eon_msm_srs", want_sc.reshape(1, 2 * cols, 4), 1, 2 * cols, 2 * cols, want)
        ok = bool(np.array_equal(want.reshape(2, cols, 8), o["wits"]))
        parity["open_witness_identity"] = H.all_ok(ok)
        open_ms = o["ms"] / args.steps
        adds_open = rows * 2 * cols * W
        open_obj = {"metric": "kzg_open_cols_rows_per_s", "value": rows * cols_total * args.steps / (o["ms"] * 1e-3),
                    "unit": UNIT, "ms_per_step": open_ms, "points": 2, "cols": cols_total,
                    "what": "eon_kzg_open: 2 points x this rank's columns -> opened values + witnesses (host buffers in "
                            "and out, so this is also the e2e number); one batched MSM over all (point, column) pairs",
                    "h2d_bytes_per_step": 64, "d2h_bytes_per_step": 2 * cols * 96, "gpu_launches": o["launches"],
                    "phase_ms_per_step": {k: v / args.steps for k, v in o["phases"].items()}}
        if rank == 0:
            open_obj["roofline"] = imad_roofline(ctx, o["phases"], args.steps, adds_open, o["rounds"], "open",
                                                 "open_tree_bwd")
    ctx.call("eon_handle_free", h0)

    # ---- weak scaling side number: every rank its own `cols_total` columns ----------------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        d_ev_w = synth_device(rows, cols_total, 100 * (rank + 1), H.dev)
        d_lde_w = torch.empty((rows << ab, cols_total, 4), dtype=torch.int64, device=H.dev)
        cm_w = np.zeros((cols_total, 8), dtype=np.uint64)

        def step_weak():
            h = C.c_uint64(0)
            ctx.call("eon_kzg_commit_lde_dev", C.c_void_p(d_ev_w.data_ptr()), log_rows, cols_total, shift_one, cm_w,
                     C.byref(h), log_rows + ab, shift_lde, C.c_void_p(d_lde_w.data_ptr()))
            ctx.call("eon_handle_free", h)
        for _ in range(args.warmup):
            step_weak()
        ms_w = H.timed(step_weak, args.steps)
        weak = {"value": rows * cols_total * world * args.steps / (ms_w * 1e-3), "unit": UNIT,
                "ms_per_step": ms_w / args.steps, "what": f"every rank its own 2^{log_rows} x {cols_total} trace"}
        del d_ev_w, d_lde_w

    # ---- concurrent PCIe rates (what bounds the host-buffer leg at N ranks) -----------------------------------------
    pcie = None
    if not args.no_e2e:
        probe_rows = min(rows, 1 << 19)
        ms_h2d, ms_d2h = C.c_float(), C.c_float()
        H.barrier()
        ctx.call("eon_bench_copy2d", C.c_void_p(lde_pin_np.ctypes.data), probe_rows, cols_total * 32, cols_total * 32, 1,
                 C.byref(ms_h2d))
        H.barrier()
        ctx.call("eon_bench_copy2d", C.c_void_p(lde_pin_np.ctypes.data), probe_rows, cols_total * 32, cols_total * 32, 0,
                 C.byref(ms_d2h))
        H.barrier()
        by = probe_rows * cols_total * 32
        rates = H.gather_u64(np.array([int(by / ms_h2d.value / 1e3), int(by / ms_d2h.value / 1e3)], dtype=np.uint64))
        pcie = {"what": f"contiguous {by >> 20} MiB pinned<->device copy, all {world} ranks at the same time, MB/s per rank",
                "h2d_mb_s": [int(r[0]) for r in rates], "d2h_mb_s": [int(r[1]) for r in rates]}

    # ---- one process, all N GPUs: the multi-device context (rank 0 alone; the others wait) --------------------------
    mctx_obj = None
    if not args.no_e2e and not args.no_mctx:
        H.host_barrier()
        if rank == 0:
            ndev = max(world, args.mctx_devices or 0) if world > 1 else (args.mctx_devices or 1)
            ndev = min(ndev, torch.cuda.device_count())
            try:
                m = eon.MultiContext(list(range(ndev)))
                m.call("eon_srs_generate_unsafe", field.to_wire(ALPHA), rows)
                cm = np.zeros((cols_total, 8), dtype=np.uint64)

                def step_mctx():
                    h = C.c_uint64(0)
                    m.call("eon_kzg_commit_lde", host_pin_np, log_rows, cols_total, shift_one, cm, C.byref(h),
                           log_rows + ab, shift_lde, lde_pin_np)
                    m.call("eon_handle_free", h)
                for _ in range(max(1, args.warmup)):
                    step_mctx()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    step_mctx()
                dt = time.perf_counter() - t0
                ok_m = bool(np.array_equal(cm[c0:c1], commits_first))   # rank 0's own columns of the whole commitment
                mctx_obj = {"devices": ndev, "ms_per_step": dt / args.steps * 1e3,
                            "value": rows * cols_total * args.steps / dt, "unit": UNIT,
                            "call": "eon_mctx_kzg_commit_lde on the whole host matrix (one process, one thread per GPU, "
                                    "columns sharded inside the library); wall clock around synchronous calls",
                            "commit_matches_single_context": ok_m}
                m.close()
            except Exception as e:  # report, do not lose the main line
                mctx_obj = {"error": str(e)[:300]}
        H.host_barrier()

    # ---- standalone MSM 2^24 (BASELINE configs[2], the metric's second half) ----------------------------------------
    msm_obj = None
    if args.msm_log_n > 0:
        del d_evals, d_lde
        torch.cuda.empty_cache()
        msm_obj, msm_phases = msm_leg(H, ctx, args.msm_log_n, 1, args.steps, args.warmup, C)
        if rank == 0:
            nsh = (1 << args.msm_log_n) // world
            c_used = msm_obj["msm_window_bits"]
            msm_obj["roofline"] = imad_roofline(ctx, msm_phases, args.steps, nsh * ((255 + c_used - 1) // c_used),
                                                msm_obj["msm_affine_rounds"], f"msm 2^{args.msm_log_n} shard",
                                                "msm_2p24_tree_bwd")
        parity["msm_eq_dlog_sum"] = msm_obj["parity_ok"]

    parity_ok = all(parity.values())
    if rank != 0:
        H.done()
        if not parity_ok:
            raise SystemExit(3)
        return

    # ---- rooflines (per-kernel durations from the two-call run, where the same kernels run back to back) ------------
    adds = rows * cols * W                                              # bucket additions of this rank's shard
    roofline = imad_roofline(ctx, phases, args.steps, adds, rounds_commit, "commit", "commit_tree_bwd")
    hbm_peak, peak_src = peaks()
    ntt_ms = phases["ntt_passes"] / args.steps
    ntt_bytes = 3 * 2 * (rows * cols * 32) + 3 * 2 * ((rows << ab) * cols * 32) - ((rows << ab) - rows) * cols * 32
    imad_bf = 214 if int(ctx.lib.eon_ntt_twiddle_form()) == 1 else IMAD_PER_MODMUL
    ntt_traffic, ntt_tsrc = ncu_traffic("commit_ntt_passes")
    roofline_ntt = {
        "kernel": "k_ntt_pass (HBM passes of the 2^%d iDFT and of the 2^%d LDE)" % (log_rows, log_rows + ab),
        "bound": "hbm", "achieved": ntt_bytes / (ntt_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": ntt_bytes / (ntt_ms * 1e-3) / 1e9 / hbm_peak, "traffic": ntt_traffic, "traffic_source": ntt_tsrc,
        "launch_ms_total": ntt_ms, "peak_source": peak_src, "imad_per_butterfly": imad_bf,
        "imad_frac": (((rows // 2) * log_rows * cols + ((rows << ab) // 2) * log_rows * cols) * imad_bf
                      + rows * cols * IMAD_PER_MODMUL) / (ntt_ms * 1e-3) / 1e12 / roofline["peak"],
    }

    # ---- CPU baseline: the port on the host cores (N = 1 only) ---------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu and not args.no_e2e:
        os.environ.setdefault("OMP_NUM_THREADS", str(len(os.sched_getaffinity(0))))
        from oracle import cport
        srs_host = cport.srs_generate(ALPHA, rows)
        sample_cols = min(cols_total, 4)
        est, d = cpu_reference_step(log_rows, cols_total, ab, srs_host, host_pin_np, msm_cols=sample_cols)
        cpu = {"value": rows * cols_total / est, "unit": UNIT, "cores": cport.num_threads(), "kind": "port",
               "sample": f"one step of the same workload: full coset iDFT + LDE of 2^{log_rows}x{cols_total}; MSM on "
                         f"{sample_cols}/{cols_total} columns scaled (the reference arm, --impl reference, runs all of "
                         f"them); est. {est:.2f} s per step", "detail": d}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": shared_config(args),
        "config_detail": {"parallelism": f"columns x{world}: rank r owns column_shard({cols_total}, {world}, r) of the SAME "
                                         f"trace ({cols} columns on rank 0)",
                          "call": "one eon_kzg_commit_lde_dev per step and rank (LDE on a second stream beside the MSM), "
                                  "then an all_gather of the commitments",
                          "msm_window_bits": c_bits, "msm_windows": W,
                          "l2": "inputs larger than L2: per rank %d MiB trace + %d MiB LDE + sort / pair workspace of "
                                "several hundred MiB against the 126 MB L2" % (rows * cols * 32 >> 20,
                                                                              (rows << ab) * cols * 32 >> 20),
                          "host_affinity": (f"rank 0 bound to {len(H.cpus)} cores next to its GPU (NVML)" if H.cpus else "none")},
        "parity_ok": parity_ok, "parity": parity,
        "e2e": None if args.no_e2e else {
            "value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
            "h2d_bytes_per_step": rows * cols_total * 32, "d2h_bytes_per_step": (rows << ab) * cols_total * 32 + cols_total * 64,
            "bytes_are": "whole job (all ranks); every rank moves its columns of the one host matrix",
            "call": ("plonky3_eon_b200.dist.RowBlockCommitLde: contiguous row blocks over PCIe, all_to_all over NVLink, "
                     "eon_coset_lde_batch_dev + eon_kzg_commit_dev on this rank's columns"
                     if (ms_rowblock is not None and ms_rowblock == ms_e2e) else
                     "eon_kzg_commit_lde_ld (Pcs::commit with an LDE hint) on this rank's columns of the pinned host "
                     "matrix, LDE columns written back into the pinned host result"),
            "strided_ms_per_step": ms_e2e_strided / args.steps,
            "phase_ms_per_step": {k: v / args.steps for k, v in phases_e2e.items()},
            "rowblock_ms_per_step": (ms_rowblock / args.steps) if ms_rowblock is not None else None,
            "two_calls_ms_per_step": ms_e2e2 / args.steps, "two_calls_value": units / (ms_e2e2 * 1e-3),
            "two_calls": "eon_kzg_commit_ld then eon_kzg_evals_on_coset_ld (unhinted Pcs::commit + "
                         "get_evaluations_on_domain)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "two_calls": {"ms_per_step": ms_dev2 / args.steps, "value": units / (ms_dev2 * 1e-3),
                      "what": "eon_kzg_commit_dev then eon_kzg_evals_on_coset_dev (LDE after the MSM, one stream)"},
        "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
        "phase_ms_per_step_fused": {k: v / args.steps for k, v in phases_fused.items()},
        "roofline": roofline, "roofline_ntt": roofline_ntt, "cpu_baseline": cpu,
        # msm_tree_* and msm_finish are sub-intervals of msm_accumulate (include/eon_kzg.h): not added twice
        "msm_points_per_s": rows * cols / (sum(phases[k] for k in ("msm_digits", "msm_scan", "msm_scatter",
                                                                   "msm_accumulate", "msm_reduce")) / args.steps * 1e-3),
        "open": open_obj, "msm_2p24": msm_obj, "weak": weak, "mctx": mctx_obj, "pcie_concurrent": pcie,
    }
    print(json.dumps(line), file=OUT, flush=True)
    H.done()
    if not parity_ok:
        raise SystemExit(3)


# ------------------------------------------------------------------------------------------------
def run_msm(args):
    """BASELINE configs[2] as its own line: standalone G1 Pippenger MSM, 2^log_n points x `cols` columns."""
    import ctypes as C

    import plonky3_eon_b200 as eon
    H = Harness()
    ctx = eon.Context(H.local, stream=H.stream.cuda_stream)
    if args.slice_schedule >= 0:
        ctx.call("eon_msm_set_slice_schedule", args.slice_schedule)
    if args.msm_rounds >= 0:
        ctx.call("eon_msm_set_rounds", args.msm_rounds)
    sampler = ClockSampler(H.local)
    global MSM_WINDOW_BITS
    MSM_WINDOW_BITS = args.window_bits
    if H.rank == 0:
        sampler.start()
    res, phases = msm_leg(H, ctx, args.log_n, args.msm_cols, args.steps, args.warmup, C)
    clocks = sampler.stop() if H.rank == 0 else None
    if H.rank == 0:
        n = 1 << args.log_n
        line = dict(res)
        line.update({"steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                     "dtype": DTYPE, "data": "synthetic", "clocks": clocks,
                     "config": {"workload": f"standalone G1 MSM, 2^{args.log_n} points x {args.msm_cols} scalar column(s), "
                                            f"random Fr scalars (BASELINE configs[2]); point-index sharded over "
                                            f"{H.world} GPU(s)", "points": n, "cols": args.msm_cols}})
        print(json.dumps(line), file=OUT, flush=True)
    H.done()


def run_prove_pcs(args):
    """BASELINE configs[3], PCS side only: the reference has no BN254 Poseidon2 AIR, so this
    replays the PCS call sequence of eon_uni_stark::prove (eon-uni-stark/src/prover.rs:186-187, 307-322,
    371-372, 416-442) on a random 2^log_rows x cols trace through the host mirror of the Pcs trait:
    commit(trace) -> get_evaluations_on_domain(quotient coset) -> commit_quotient(2 chunks) ->
    open(trace at zeta and zeta*omega, chunks at zeta).  Host buffers throughout; wall-clock per phase.
    --mctx-devices N runs the same sequence through ONE multi-device context over N GPUs."""
    import torch

    import plonky3_eon_b200 as eon

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the eon arm has no CPU fallback")
    rows, cols, log_rows = 1 << args.log_rows, args.cols, args.log_rows
    ndev = args.mctx_devices or 1
    ctx = eon.MultiContext(list(range(ndev))) if ndev > 1 else eon.Context(0)
    pcs = eon.GpuKzgPcs.new(rows - 1, ALPHA, ctx=ctx).with_lde_hint(1)
    trace = eon.pinned_empty((rows, cols, 4))        # the prover keeps its matrices in page-locked memory
    trace[:] = synth_trace(rows, cols)
    quotient = eon.pinned_empty((2 * rows, 1, 4))    # stands in for quotient_values (CPU side in the reference)
    quotient[:] = synth_trace(2 * rows, 1, c0=500)
    dom = pcs.natural_domain_for_degree(rows)
    qdom = dom.create_disjoint_domain(2 * rows)
    zeta = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF

    def once():
        t = [time.perf_counter()]
        tc, tpd = pcs.commit([(dom, trace)])
        t.append(time.perf_counter())
        lde = pcs.get_evaluations_on_domain(tpd, 0, qdom)
        t.append(time.perf_counter())
        qc, qpd = pcs.commit_quotient(qdom, quotient, 2)
        t.append(time.perf_counter())
        opened, proof = pcs.open([(tpd, [[zeta, dom.next_point(zeta)]]), (qpd, [[zeta], [zeta]])])
        t.append(time.perf_counter())
        for m in tpd + qpd:
            m.free()
        assert lde.shape[0] == 2 * rows and len(proof) == 2
        return [b - a for a, b in zip(t, t[1:])]

    for _ in range(args.warmup):
        once()
    launches0 = ctx.launch_count()
    runs = [once() for _ in range(args.steps)]
    launches = ctx.launch_count() - launches0
    med = [float(np.median([r[i] for r in runs])) for i in range(4)]
    total = sum(med)
    line = {
        "metric": "prove_pcs_cols_rows_per_s", "value": rows * cols / total, "unit": "cols*rows/s", "n_gpus": ndev,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": f"PCS call sequence of eon_uni_stark::prove on a random 2^{log_rows} x {cols} trace, "
                               "log_quotient_degree 1, 2 quotient chunks, page-locked host buffers, Python mirror of the "
                               "Pcs trait (BASELINE configs[3] without the AIR)",
                   "rows": rows, "cols": cols,
                   "context": f"one eon_mctx over {ndev} GPUs" if ndev > 1 else "one eon_ctx"},
        "phase_ms": {"commit_trace_with_lde_hint": med[0] * 1e3, "get_evaluations_on_domain": med[1] * 1e3,
                     "commit_quotient_2_chunks": med[2] * 1e3, "open": med[3] * 1e3},
        "gpu_launches": launches,
    }
    print(json.dumps(line), file=OUT, flush=True)


OUT = sys.stdout


def claim_stdout():
    """Keep stdout to the one JSON line: fd 1 is pointed at stderr for the life of the process (native libraries
    print there -- NCCL's version banner did) and the line is written to a duplicate of the real stdout."""
    global OUT
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    OUT = os.fdopen(real, "w")


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="eon", choices=["eon", "reference"])
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--cols", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--added-bits", type=int, default=1, help="log2 of the LDE blow-up (1: configs[1]; 2: configs[4])")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (large shapes: the trace is then "
                                                          "generated on the device, per rank shard)")
    ap.add_argument("--no-open", action="store_true", help="skip the KzgPcs::open leg")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling side number")
    ap.add_argument("--no-mctx", action="store_true", help="skip the one-process multi-device-context leg")
    ap.add_argument("--mctx-devices", type=int, default=0, help="GPUs the multi-device-context leg drives (default: N)")
    ap.add_argument("--rowblock-max-cols", type=int, default=4,
                    help="N > 1: columns per rank up to which the e2e leg also tries the row-block transport")
    ap.add_argument("--msm-log-n", type=int, default=24, help="standalone MSM side leg: log2 points (0 = skip)")
    ap.add_argument("--check-e2e", action="store_true", help="also compare the fused and two-call device LDE bytes")
    ap.add_argument("--warmup-ref", type=int, default=0,
                    help="reference arm: warm-up steps if they should differ from --warmup (0 = use --warmup)")
    ap.add_argument("--window-bits", type=int, default=-1,
                    help="MSM window tables: -1 library default, 0 none (plain c=16), 8..20 explicit")
    ap.add_argument("--slice-schedule", type=int, default=-1,
                    help="MSM round 0 walked by table slice: 1 on, 0 off, -1 the library's policy")
    ap.add_argument("--msm-rounds", type=int, default=-1,
                    help="batched-affine pairwise rounds before the XYZZ finisher: -1 the library's policy, 0..6 explicit")
    ap.add_argument("--workload", default="commit", choices=["commit", "msm", "prove-pcs"],
                    help="commit: KZG commit + LDE (the headline metric, with open / MSM 2^24 side legs); msm: standalone "
                         "MSM (configs[2]); prove-pcs: the PCS call sequence of a proof")
    ap.add_argument("--log-n", type=int, default=24, help="msm workload: log2 of the point count")
    ap.add_argument("--msm-cols", type=int, default=1, help="msm workload: scalar columns sharing the bases")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "msm":
        run_msm(args)
    elif args.workload == "prove-pcs":
        run_prove_pcs(args)
    else:
        run_eon(args)


if __name__ == "__main__":
    main()
