#!/usr/bin/env python
"""bench.py — BN254 KZG commit + coset LDE throughput (BASELINE.json metric) on N B200s.

One "step" = the hot path over one synthetic trace of 2^20 rows x 16 columns per GPU
(BASELINE.json configs[1]):
    KzgPcs::commit                 = coset iDFT (2^20 x 16) + 16 G1 MSMs of 2^20 points   (kzg/src/pcs.rs:223-265)
    get_evaluations_on_domain      = zero-pad + coset NTT onto the 2^21-point coset 5*K      (blow-up 2)
value = cols*rows / s with the trace resident in HBM (CUDA events on the launch stream).
e2e   = the same through the host-buffer C ABI: pinned host trace -> H2D -> commit -> LDE -> D2H.
N > 1: column sharding, every rank owns its own 2^20 x 16 slab of a 2^20 x (16 N) trace ("weak"); the
only exchange is an all_gather of the 16 commitments (1 KiB) per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl eon|reference] [--log-rows 20] [--cols 16]

--impl reference times the CPU port of the reference path (oracle/c, OpenMP, all host threads) on a
bounded sample of the same workload; the reference itself is Rust + halo2curves and cannot be built here.
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


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev = dev
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # under load = samples at or above the median
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synth_trace(seed, rows, cols):
    """Uniform Fr in Montgomery form exactly like the reference sampler (bn254/src/field.rs:534-551):
    random 256 bits, top 2 bits cleared, rejected if >= P, used AS the Montgomery limbs."""
    from plonky3_eon_b200 import field
    rng = np.random.default_rng(seed)
    n = rows * cols
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 62) - 1)
    pl = [(field.P >> (64 * i)) & ((1 << 64) - 1) for i in range(4)]
    while True:
        lt = np.zeros(n, dtype=bool)
        eq = np.ones(n, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (out[:, k] < np.uint64(pl[k]))
            eq &= out[:, k] == np.uint64(pl[k])
        bad = np.nonzero(~lt)[0]
        if len(bad) == 0:
            break
        rep = rng.integers(0, 1 << 64, size=(len(bad), 4), dtype=np.uint64)
        rep[:, 3] &= np.uint64((1 << 62) - 1)
        out[bad] = rep
    return out.reshape(rows, cols, 4)


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPU cores next to its GPU before any pinned host buffer is allocated
    (first touch then places the buffers on that NUMA node): with 8 ranks the host<->device copies of
    the e2e leg otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(log_rows, cols, sample_cols, srs=None, repeats=1):
    """The CPU port (oracle/c) on a bounded sample: full iDFT + LDE of the rows x cols trace, MSM on
    `sample_cols` of the columns (scaled to `cols`).  Returns (value, seconds_estimated, detail)."""
    from oracle import cport
    rows = 1 << log_rows
    ev = synth_trace(1, rows, cols)
    if srs is None:
        srs = cport.srs_generate(ALPHA, rows)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        commits, coeffs = cport.kzg_commit(ev, 1, srs, ncols_msm=0)           # coset iDFT only
        t1 = time.perf_counter()
        cport.msm(srs, coeffs, ncols=sample_cols, ld=cols)                       # sample of the MSMs
        t2 = time.perf_counter()
        pad = np.zeros((2 * rows, cols, 4), dtype=np.uint64)
        pad[:rows] = coeffs
        t3 = time.perf_counter()
        cport.coset_dft_batch(pad, SHIFT_LDE)                                    # zero-pad + coset DFT (2^21)
        t4 = time.perf_counter()
        est = (t1 - t0) + (t2 - t1) * (cols / sample_cols) + (t4 - t3)
        d = {"idft_s": t1 - t0, "msm_sample_s": t2 - t1, "lde_s": t4 - t3}
        if best is None or est < best[0]:
            best = (est, d)
    est, d = best
    return rows * cols / est, est, d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is rank 0 alone and is meant to use
    # every host core it can get (the reference's rayon pool would), so undo that before OpenMP starts
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    from oracle import cport
    cores = cport.num_threads()
    sample_cols = 2 if args.cols >= 2 else 1
    rows = 1 << args.log_rows
    t_setup = time.perf_counter()
    srs = cport.srs_generate(ALPHA, rows)
    t_setup = time.perf_counter() - t_setup
    vals = []
    # exactly --warmup untimed and --steps timed steps, like the eon arm; every step is the bounded sample above
    # (about 3 s of CPU work at 2^20 x 16 on the GPU box's host), so the default run stays well under a few minutes
    steps = max(1, args.steps)
    warm = args.warmup_ref if args.warmup_ref > 0 else max(0, args.warmup)
    for i in range(warm + steps):
        v, est, d = cpu_reference_sample(args.log_rows, args.cols, sample_cols, srs=srs)
        if i >= warm:
            vals.append((v, est, d))
    v = float(np.median([x[0] for x in vals]))
    est = float(np.median([x[1] for x in vals]))
    sample = (f"full coset iDFT 2^{args.log_rows}x{args.cols} + coset LDE to 2^{args.log_rows + 1} rows on all "
              f"{args.cols} columns; MSM on {sample_cols} of {args.cols} columns, scaled x{args.cols // sample_cols}; "
              f"affine SRS normalised once (no per-call to_affine, bn254/src/curve.rs:170)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery, BN254 Fr/Fq)", "data": "synthetic",
        "config": {"workload": f"KZG commit + blow-up-2 coset LDE, 2^{args.log_rows} rows x {args.cols} cols, "
                               "CPU port of the reference path (oracle/c, OpenMP)",
                   "srs_setup_s": t_setup},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "detail": vals[-1][2]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ------------------------------------------------------------------------------------------------
def run_eon(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import plonky3_eon_b200 as eon
    from plonky3_eon_b200 import field

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the eon arm has no CPU fallback")
    torch.cuda.set_device(local)
    cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    rows, cols, log_rows = 1 << args.log_rows, args.cols, args.log_rows

    stream = torch.cuda.current_stream()
    ctx = eon.Context(local, stream=stream.cuda_stream)
    pcs = eon.GpuKzgPcs.new(rows - 1, ALPHA, ctx=ctx)           # synthetic SRS alpha^i * G on the device
    if args.window_bits >= 0:                                    # default: the library's own table policy
        ctx.call("eon_srs_set_window_tables", args.window_bits)
    if args.slice_schedule >= 0:                                 # default: the library's own policy
        ctx.call("eon_msm_set_slice_schedule", args.slice_schedule)
    if args.msm_rounds >= 0:                                     # default: the library's own policy (3 rounds)
        ctx.call("eon_msm_set_rounds", args.msm_rounds)
    shift_one = field.to_wire(1)
    shift_lde = field.to_wire(SHIFT_LDE)

    # synthetic trace, pinned on the host and resident on the device
    host_np = synth_trace(1 + rank, rows, cols)
    host_pin = torch.from_numpy(host_np.view(np.int64)).pin_memory()
    host_pin_np = host_pin.numpy().view(np.uint64)
    d_evals = host_pin.to("cuda", non_blocking=False)
    ab = args.added_bits                                         # blow-up 2^ab (1 = configs[1], 2 = configs[4])
    d_lde = torch.empty((rows << ab, cols, 4), dtype=torch.int64, device="cuda")
    if args.no_e2e:
        lde_pin = lde_pin_np = None
    else:
        lde_pin = torch.empty((rows << ab, cols, 4), dtype=torch.int64).pin_memory()
        lde_pin_np = lde_pin.numpy().view(np.uint64)
    commits = np.zeros((cols, 8), dtype=np.uint64)
    gathered = [torch.empty(cols * 8, dtype=torch.int64, device="cuda") for _ in range(world)] if world > 1 else None

    def step_device():
        # commit + the hinted quotient-coset LDE in one call (the LDE transform runs beside the MSM)
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_lde_dev", C.c_void_p(d_evals.data_ptr()), log_rows, cols, shift_one, commits,
                 C.byref(h), log_rows + ab, shift_lde, C.c_void_p(d_lde.data_ptr()))
        ctx.call("eon_handle_free", h)
        if world > 1:
            t = torch.from_numpy(commits.view(np.int64).reshape(-1)).cuda()
            dist.all_gather(gathered, t)

    def step_device_two_calls():
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_dev", C.c_void_p(d_evals.data_ptr()), log_rows, cols, shift_one, commits,
                 C.byref(h))
        ctx.call("eon_kzg_evals_on_coset_dev", h, log_rows + ab, shift_lde, C.c_void_p(d_lde.data_ptr()))
        ctx.call("eon_handle_free", h)
        if world > 1:
            t = torch.from_numpy(commits.view(np.int64).reshape(-1)).cuda()
            dist.all_gather(gathered, t)

    def step_e2e():
        # what a Pcs shim with an LDE hint calls from commit() (GpuKzgPcs.with_lde_hint): commit + the
        # quotient-coset evaluations in one call, column groups pipelined over PCIe
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit_lde", host_pin_np, log_rows, cols, shift_one, commits, C.byref(h), log_rows + ab,
                 shift_lde, lde_pin_np)
        ctx.call("eon_handle_free", h)
        if world > 1:
            t = torch.from_numpy(commits.view(np.int64).reshape(-1)).cuda()
            dist.all_gather(gathered, t)

    def step_e2e_two_calls():
        # the unhinted trait sequence: Pcs::commit, then Pcs::get_evaluations_on_domain
        h = C.c_uint64(0)
        ctx.call("eon_kzg_commit", host_pin_np, log_rows, cols, shift_one, commits, C.byref(h))
        ctx.call("eon_kzg_evals_on_coset", h, log_rows + ab, shift_lde, lde_pin_np)
        ctx.call("eon_handle_free", h)
        if world > 1:
            t = torch.from_numpy(commits.view(np.int64).reshape(-1)).cuda()
            dist.all_gather(gathered, t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    ctx.phase_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ms_dev = timed(step_device, args.steps)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    phases = ctx.phase_ms()          # summed over the timed steps
    commits_device = commits.copy()
    lde_dev_fused = d_lde.clone() if args.check_e2e else None
    step_device_two_calls()
    ctx.phase_reset()
    ms_dev2 = timed(step_device_two_calls, args.steps)
    phases_serial = ctx.phase_ms()   # the same kernels run back to back: per-kernel times for the rooflines
    assert np.array_equal(commits, commits_device), "fused and two-call commitments differ"
    if lde_dev_fused is not None:
        assert torch.equal(lde_dev_fused, d_lde), "fused and two-call device LDE differ"
        del lde_dev_fused

    if args.no_e2e:
        ms_e2e = ms_e2e2 = None
    else:
        for _ in range(max(1, args.warmup)):                    # the same W untimed steps as the device leg
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        assert np.array_equal(commits, commits_device), "e2e and device-resident commitments differ"
        lde_fused = lde_pin_np.copy() if rank == 0 and args.check_e2e else None
        step_e2e_two_calls()
        ms_e2e2 = timed(step_e2e_two_calls, args.steps)
        assert np.array_equal(commits, commits_device), "two-call e2e and device-resident commitments differ"
        if lde_fused is not None:
            assert np.array_equal(lde_fused, lde_pin_np), "fused and two-call LDE differ"

    units = rows * cols * n_gpus * args.steps
    value = units / (ms_dev * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3) if ms_e2e else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # per-kernel durations for the rooflines come from the two-call timed run, where the same kernels run
    # back to back on one stream (in the fused step the LDE passes share the GPU with the MSM kernels)
    phases_overlapped = phases
    phases = phases_serial

    # ---- roofline of the dominant kernel (MSM bucket accumulation: integer pipe) -----------------
    imad_peak = max(ctx.imad_peak_tops(0), ctx.imad_peak_tops(1))       # T IMAD/s, measured now
    imad_wide = ctx.imad_peak_tops(2)
    modmul_g = ctx.modmul_gmuls(1)
    c_bits = int(ctx.lib.eon_srs_window_bits(ctx.h)) or 16              # window tables in use (0 = plain c = 16)
    W = (255 + c_bits - 1) // c_bits                                    # msm_windows() in csrc/msm.cu
    adds = rows * cols * W                                              # one bucket addition per (point, window, column)
    rounds = int(ctx.lib.eon_msm_rounds_used(ctx.h))
    acc_ms = phases["msm_accumulate"] / args.steps
    hbm_peak, peak_src = peaks()
    ntt_ms = phases["ntt_passes"] / args.steps
    ntt_bytes = 3 * 2 * (rows * cols * 32) + 3 * 2 * ((rows << ab) * cols * 32) - ((rows << ab) - rows) * cols * 32
    common = {
        "bound": "imad", "peak": imad_peak, "unit": "TIMAD/s", "traffic": None,
        "peak_source": "eon_bench_imad_peak in this run (mad.lo/mad.hi.u32, 16 independent chains/thread)",
        "note": "the dominant kernel is integer-pipe bound (SURVEY §8d), so the roofline is IMAD, not HBM/tensor; "
                "mad.wide peak (counted as 2 ops) and measured Fq modmul rate are given beside it",
        "imad_wide_tops": imad_wide, "fq_modmul_gmul_s": modmul_g,
        "modmul_equiv_tops": modmul_g * 1e9 * IMAD_PER_MODMUL / 1e12,
    }
    if rounds:
        # batched-affine pairwise rounds (csrc/msm_tree.cu): k_tree_bwd does 5 Fq products per pair
        # (2 to peel the shared inverse, lambda, lambda^2, y3); round r has adds / 2^(r+1) pairs
        pairs = sum(adds // (1 << (r + 1)) for r in range(rounds))
        bwd_ms = phases["msm_tree_bwd"] / args.steps
        ops = pairs * 5 * IMAD_PER_MODMUL
        achieved = ops / (bwd_ms * 1e-3) / 1e12
        # HBM view of the same kernel family: per pair 2 points in (128 B), prefix product in (32 B), sum out (64 B)
        alg_bytes = pairs * 224
        # ncu dram__bytes_read/write per launch at 2^20 x 16 (profiles/r01k_ncu_launches.csv).  With round 0 walked
        # by table slice (csrc/msm_tree.cu) its launch moves 10.89 GB read + 8.25 GB written for 125.8 M pairs =
        # 152 B per pair (the base gathers hit the L2; before: 40.25 + 8.28 GB = 386 B per pair, DRAM fetching 128 B
        # per random 64-byte gather); the dense rounds move 14.64 + 7.32 GB for 94.4 M pairs = 233 B per pair.
        traffic = None
        if log_rows == 20 and cols == 16 and rounds == 3:
            r0 = (40.25e9 + 8.28e9) if args.slice_schedule == 0 else (10.89e9 + 8.25e9)
            traffic = r0 + 10.52e9 + 4.12e9 + 5.27e9 + 2.05e9
        roofline = dict(common, kernel=f"k_tree_bwd x{rounds} (batched-affine pair additions, 5 modmul per pair)",
                        achieved=achieved, frac=achieved / imad_peak, algorithmic_ops_per_launch=ops,
                        launch_ms=bwd_ms, traffic=traffic,
                        traffic_source="ncu dram__bytes_read.sum + dram__bytes_write.sum of these launches "
                                       "(profiles/r01k_ncu_launches.csv), sum over the 3 rounds of one step"
                        if traffic else None,
                        hbm_view={"bound": "hbm", "algorithmic_bytes": alg_bytes,
                                  "achieved": alg_bytes / (bwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": alg_bytes / (bwd_ms * 1e-3) / 1e9 / hbm_peak},
                        accumulate_phase={"ms": acc_ms, "tree_fwd_ms": phases["msm_tree_fwd"] / args.steps,
                                          "tree_inv_ms": phases["msm_tree_inv"] / args.steps,
                                          "tree_bwd_ms": bwd_ms, "finish_ms": phases["msm_finish"] / args.steps,
                                          "bucket_additions": adds,
                                          "xyzz_equiv_frac": adds * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL
                                          / (acc_ms * 1e-3) / 1e12 / imad_peak})
    else:
        achieved = adds * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL / (acc_ms * 1e-3) / 1e12
        roofline = dict(common, kernel="k_msm_accumulate (XYZZ mixed adds, one thread per bucket)",
                        achieved=achieved, frac=achieved / imad_peak,
                        algorithmic_ops_per_launch=adds * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL, launch_ms=acc_ms)
    # fixed-operand (Shoup) twiddle product: 43 + 36 + 36 limb products = 214 IMAD (csrc/fp_shoup.cuh)
    imad_bf = 214 if int(ctx.lib.eon_ntt_twiddle_form()) == 1 else IMAD_PER_MODMUL
    roofline_ntt = {
        "kernel": "k_ntt_pass (3 HBM passes for the 2^20 iDFT + 3 for the 2^21 LDE)",
        "bound": "hbm", "achieved": ntt_bytes / (ntt_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": ntt_bytes / (ntt_ms * 1e-3) / 1e9 / hbm_peak,
        # ncu --set full (profiles/r01h_ncu_full_summary.txt): 1.05-1.07 GB per 2^20 pass, 1.57-2.1 GB per 2^21 pass
        "traffic": (3.20e9 + 5.76e9) if (log_rows == 20 and cols == 16) else None, "launch_ms_total": ntt_ms,
        "peak_source": peak_src,
        # butterflies of the iDFT and of the LDE (its first `ab` layers are replication) at the IMAD count of the
        # product in use, plus the 1/n Montgomery product on every iDFT output
        "imad_per_butterfly": imad_bf,
        "imad_frac": (((rows // 2) * log_rows * cols + ((rows << ab) // 2) * log_rows * cols) * imad_bf
                      + rows * cols * IMAD_PER_MODMUL) / (ntt_ms * 1e-3) / 1e12 / imad_peak,
    }

    # ---- CPU baseline: the port on the host cores, bounded sample (N = 1 only) -------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import cport
        srs_host = pcs.g1_powers()                      # same SRS, read back (setup, untimed)
        sample_cols = 2 if cols >= 2 else 1
        v, est, d = cpu_reference_sample(log_rows, cols, sample_cols, srs=srs_host)
        cpu = {"value": v, "unit": UNIT, "cores": cport.num_threads(), "kind": "port",
               "sample": f"full coset iDFT + LDE of 2^{log_rows}x{cols}; MSM on {sample_cols}/{cols} columns scaled; "
                         f"est. {est:.2f} s per step", "detail": d}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (8x32-bit Montgomery limbs, BN254 Fr/Fq)", "data": "synthetic",
        "config": {"workload": f"KZG commit (coset iDFT + {cols} MSM) + blow-up-{1 << ab} coset LDE, 2^{log_rows} rows x "
                               f"{cols} cols per GPU (BASELINE configs[1]); column-sharded {cols * n_gpus} cols total; "
                               "one eon_kzg_commit_lde call per step (LDE on a second stream beside the MSM)",
                   "rows": rows, "cols_per_gpu": cols, "srs_points": rows, "msm_window_bits": c_bits, "msm_windows": W,
                   "msm_affine_rounds": rounds,
                   "l2": "inputs (512 MiB trace, 1 GiB LDE, 1 GiB sort workspace) exceed the 126 MB L2",
                   "parallelism": f"columns x{n_gpus}",
                   "host_affinity": (f"rank 0 bound to {len(cpus)} cores next to its GPU (NVML)" if cpus else "none")},
        "e2e": None if args.no_e2e else {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": rows * cols * 32, "d2h_bytes_per_step": (rows << ab) * cols * 32 + cols * 64,
                "call": "eon_kzg_commit_lde (Pcs::commit with an LDE hint; host pinned buffers in and out)",
                "two_calls_ms_per_step": ms_e2e2 / args.steps,
                "two_calls_value": units / (ms_e2e2 * 1e-3),
                "two_calls": "eon_kzg_commit then eon_kzg_evals_on_coset (unhinted Pcs::commit + "
                             "get_evaluations_on_domain)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "two_calls": {"ms_per_step": ms_dev2 / args.steps, "value": units / (ms_dev2 * 1e-3),
                      "what": "eon_kzg_commit_dev then eon_kzg_evals_on_coset_dev (LDE after the MSM, one stream)"},
        "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
        "phase_ms_per_step_fused": {k: v / args.steps for k, v in phases_overlapped.items()},
        "roofline": roofline, "roofline_ntt": roofline_ntt, "cpu_baseline": cpu,
        "msm_points_per_s": rows * cols / (sum(phases[k] for k in phases if k.startswith("msm_")) / args.steps * 1e-3),
    }
    print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_msm(args):
    """BASELINE configs[2]: standalone G1 Pippenger MSM, 2^log_n points x `cols` scalar columns with
    random Fr scalars over the synthetic SRS.  N GPUs: point-index sharding ("strong": total points
    fixed), per-rank partial sums all_gathered over NCCL and added with eon_g1_sum."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import plonky3_eon_b200 as eon
    from plonky3_eon_b200 import dist as edist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the eon arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, cols = 1 << args.log_n, args.msm_cols
    stream = torch.cuda.current_stream()
    ctx = eon.Context(local, stream=stream.cuda_stream)
    pcs = eon.GpuKzgPcs.new(n - 1, ALPHA, ctx=ctx)              # replicated SRS (every rank reads its slice)
    if args.window_bits >= 0:
        ctx.call("eon_srs_set_window_tables", args.window_bits)
    if args.slice_schedule >= 0:
        ctx.call("eon_msm_set_slice_schedule", args.slice_schedule)
    if args.msm_rounds >= 0:
        ctx.call("eon_msm_set_rounds", args.msm_rounds)
    first, cnt = edist.index_shard(n, world, rank)
    host = synth_trace(42, cnt, cols)                            # seed 42: bn254/benches/bench_curve.rs:40
    d_sc = torch.from_numpy(host.view(np.int64)).to(dev)
    out = np.zeros((cols, 8), dtype=np.uint64)
    gathered = [torch.empty(cols * 8, dtype=torch.int64, device=dev) for _ in range(world)] if world > 1 else None
    total = np.zeros((cols, 8), dtype=np.uint64)

    def step():
        ctx.call("eon_msm_srs_range_dev", C.c_void_p(d_sc.data_ptr()), first, cnt, cols, cols, out)
        if world > 1:
            t = torch.from_numpy(out.view(np.int64).reshape(-1)).to(dev)
            dist.all_gather(gathered, t)
            parts = torch.stack(gathered).cpu().numpy().view(np.uint64).reshape(world, cols, 8)
            for c in range(cols):
                ctx.call("eon_g1_sum", np.ascontiguousarray(parts[:, c]), world, total[c])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.phase_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    phases = ctx.phase_ms()
    if rank == 0:
        imad_peak = max(ctx.imad_peak_tops(0), ctx.imad_peak_tops(1))
        c_bits = int(ctx.lib.eon_srs_window_bits(ctx.h)) or min(16, max(2, (cnt - 1).bit_length() - 3))
        W = (255 + c_bits - 1) // c_bits
        acc_ms = phases["msm_accumulate"] / args.steps
        ops = cnt * cols * W * MODMUL_PER_MIXED_ADD * IMAD_PER_MODMUL
        line = {
            "metric": "msm_points_per_s", "value": n * cols * args.steps / (ms * 1e-3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u256 (8x32-bit Montgomery limbs, BN254 Fr/Fq)", "data": "synthetic",
            "config": {"workload": f"standalone G1 MSM, 2^{args.log_n} points x {cols} scalar column(s), random Fr "
                                   f"scalars (BASELINE configs[2]); point-index sharded over {world} GPU(s)",
                       "points": n, "cols": cols, "msm_window_bits": c_bits, "msm_windows": W,
                       "l2": f"scalars {cnt * cols * 32 >> 20} MiB + bases {cnt * W * 64 >> 20} MiB per GPU"},
            "gpu_launches": launches, "clocks": clocks,
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "roofline": {"kernel": "k_msm_accumulate", "bound": "imad", "achieved": ops / (acc_ms * 1e-3) / 1e12,
                         "peak": imad_peak, "unit": "TIMAD/s", "frac": ops / (acc_ms * 1e-3) / 1e12 / imad_peak,
                         "traffic": None, "algorithmic_ops_per_launch": ops, "launch_ms": acc_ms},
        }
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_open(args):
    """KzgPcs::open (kzg/src/pcs.rs:289-335) of one committed 2^log_rows x cols matrix at the two points the
    prover uses (zeta, zeta * omega; eon-uni-stark/src/prover.rs:416-431): per step 2 * cols synthetic
    divisions (k_quot_*) and one batched MSM of 2 * cols columns over the SRS.  Single GPU."""
    import ctypes as C

    import torch

    import plonky3_eon_b200 as eon
    from plonky3_eon_b200 import field

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the eon arm has no CPU fallback")
    torch.cuda.set_device(0)
    rows, cols, log_rows = 1 << args.log_rows, args.cols, args.log_rows
    stream = torch.cuda.current_stream()
    ctx = eon.Context(0, stream=stream.cuda_stream)
    eon.GpuKzgPcs.new(rows - 1, ALPHA, ctx=ctx)
    host = synth_trace(1, rows, cols)
    d_evals = torch.from_numpy(host.view(np.int64)).to("cuda")
    commits = np.zeros((cols, 8), dtype=np.uint64)
    h = C.c_uint64(0)
    ctx.call("eon_kzg_commit_dev", C.c_void_p(d_evals.data_ptr()), log_rows, cols, field.to_wire(1), commits, C.byref(h))
    zeta = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    omega = pow(field.two_adic_generator(log_rows), 1, field.P)
    pts = np.stack([field.to_wire(zeta), field.to_wire(zeta * omega % field.P)])
    vals = np.zeros((2, cols, 4), dtype=np.uint64)
    wits = np.zeros((2, cols, 8), dtype=np.uint64)

    def step():
        ctx.call("eon_kzg_open", h, pts, 2, vals, wits)

    for _ in range(args.warmup):
        step()
    ctx.phase_reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    phases = ctx.phase_ms()
    line = {
        "metric": "kzg_open_cols_rows_per_s", "value": rows * cols * args.steps / (ms * 1e-3), "unit": "cols*rows/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (8x32-bit Montgomery limbs, BN254 Fr/Fq)", "data": "synthetic",
        "config": {"workload": f"KzgPcs::open of a committed 2^{log_rows} x {cols} trace at 2 points (zeta, zeta*omega): "
                               f"{2 * cols} quotient scans + one MSM of {2 * cols} columns; values and witnesses to the host",
                   "rows": rows, "cols": cols, "points": 2},
        "gpu_launches": ctx.launch_count() - launches0,
        "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
    }
    print(json.dumps(line), file=OUT, flush=True)


def run_prove_pcs(args):
    """BASELINE configs[3], PCS side only: the reference has no BN254 Poseidon2 AIR, so this
    replays the PCS call sequence of eon_uni_stark::prove (eon-uni-stark/src/prover.rs:186-187, 307-322,
    371-372, 416-442) on a random 2^log_rows x cols trace through the host mirror of the Pcs trait:
    commit(trace) -> get_evaluations_on_domain(quotient coset) -> commit_quotient(2 chunks) ->
    open(trace at zeta and zeta*omega, chunks at zeta).  Host buffers throughout; wall-clock per phase."""
    import torch

    import plonky3_eon_b200 as eon
    from plonky3_eon_b200 import field

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the eon arm has no CPU fallback")
    rows, cols, log_rows = 1 << args.log_rows, args.cols, args.log_rows
    ctx = eon.Context(0)
    pcs = eon.GpuKzgPcs.new(rows - 1, ALPHA, ctx=ctx).with_lde_hint(1)
    trace = eon.pinned_empty((rows, cols, 4))        # the prover keeps its matrices in page-locked memory
    trace[:] = synth_trace(1, rows, cols)
    quotient = eon.pinned_empty((2 * rows, 1, 4))    # stands in for quotient_values (CPU side in the reference)
    quotient[:] = synth_trace(2, 2 * rows, 1)
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
        "metric": "prove_pcs_cols_rows_per_s", "value": rows * cols / total, "unit": "cols*rows/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs, BN254 Fr/Fq)",
        "data": "synthetic",
        "config": {"workload": f"PCS call sequence of eon_uni_stark::prove on a random 2^{log_rows} x {cols} trace, "
                               "log_quotient_degree 1, 2 quotient chunks, page-locked host buffers, Python mirror of the "
                               "Pcs trait (BASELINE configs[3] without the AIR)",
                   "rows": rows, "cols": cols},
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
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (large shapes: pinned LDE buffer)")
    ap.add_argument("--check-e2e", action="store_true", help="compare the fused and two-call LDE bytes (1 GiB copy)")
    ap.add_argument("--warmup-ref", type=int, default=0,
                    help="reference arm: warm-up steps if they should differ from --warmup (0 = use --warmup)")
    ap.add_argument("--window-bits", type=int, default=-1,
                    help="MSM window tables: -1 library default, 0 none (plain c=16), 8..20 explicit")
    ap.add_argument("--slice-schedule", type=int, default=-1,
                    help="MSM round 0 walked by 64 MiB table slice: 1 on, 0 off, -1 the library's policy")
    ap.add_argument("--msm-rounds", type=int, default=-1,
                    help="batched-affine pairwise rounds before the XYZZ finisher: -1 the library's policy, 0..6 explicit")
    ap.add_argument("--workload", default="commit", choices=["commit", "msm", "open", "prove-pcs"],
                    help="commit: KZG commit + LDE (the headline metric); msm: standalone MSM (configs[2]); "
                         "open: KzgPcs::open at 2 points")
    ap.add_argument("--log-n", type=int, default=24, help="msm workload: log2 of the point count")
    ap.add_argument("--msm-cols", type=int, default=1, help="msm workload: scalar columns sharing the bases")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "msm":
        run_msm(args)
    elif args.workload == "open":
        run_open(args)
    elif args.workload == "prove-pcs":
        run_prove_pcs(args)
    else:
        run_eon(args)


if __name__ == "__main__":
    main()
