"""bench.py — vectors quantized / second at dim=256, codebook=1024 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N=1 and per GPU for N>1): BASELINE.json configs[1] — VectorQuantize(dim=256, codebook_size=1024),
x = (64, 4096, 256) bf16, training-mode forward with the EMA codebook update.  A "step" is one such
forward over one synthetic batch.  `value` times the device-resident path with CUDA events; `e2e` times
the public module call with HOST (pinned) buffers, host<->device copies inside the timed region.
Under torchrun every rank runs the same per-GPU batch (weak scaling) with sync_codebook=True, i.e. one
NCCL all-reduce of the packed EMA statistics per step; the time is the max over ranks.

`--impl reference` times the reference's own algorithm on the host cores (the torch-CPU oracle port
oracle/vq_oracle_torch.py: (N x K) distance matrix, one-hot, three GEMMs — vector_quantize_pytorch.py:674-791)
on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vectors quantized/sec at dim=256, codebook=1024; indices bit-exact vs ref"
B, T, D, K = 64, 4096, 256, 1024
WORKLOAD = "VectorQuantize dim=256 codebook_size=1024, x=(64,4096,256) bf16, EMA on (BASELINE.json configs[1])"
CPU_SAMPLE_VECTORS = 65536  # 1/4 of the batch per CPU step (~0.7 s on 8 cores)
E2E_CHUNKS = int(os.environ.get("VQB_E2E_CHUNKS", "10"))  # row chunks of the host-buffer pipeline (forward_host)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference algorithm on the host cores)
# ------------------------------------------------------------------------------------------------

CPU_THREADS = [None]


def cpu_reference_step_factory(threads=None):
    import torch
    from oracle import vq_oracle_torch as T  # the reference's own ATen op sequence (bit-identical on the goldens)
    torch.set_num_threads(threads or os.cpu_count())
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(CPU_SAMPLE_VECTORS // 16, 16, D, generator=gen).bfloat16()
    state = T.State(torch.randn(K, D, generator=gen))

    def step():
        T.vq_forward(x, state, training=True)

    return step


def time_cpu(steps, warmup):
    """All host cores is torch's default (and what the reference would use); on many-core hosts a smaller pool is
    faster for this GEMM size, so both are timed and the FASTER one is reported (its thread count in `cores`)."""
    best = None
    for threads in sorted({os.cpu_count(), min(32, os.cpu_count())}, reverse=True):
        step = cpu_reference_step_factory(threads)
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
        cand = (CPU_SAMPLE_VECTORS * steps / dt, dt / steps * 1e3, threads)
        if best is None or cand[0] > best[0]:
            best = cand
    CPU_THREADS[0] = best[2]
    return best[0], best[1]


def cpu_baseline_block(value):
    return {"value": value, "unit": "vectors/s", "cores": CPU_THREADS[0] or os.cpu_count(), "host_cores": os.cpu_count(),
            "kind": "port",
            "sample": f"{CPU_SAMPLE_VECTORS} of the {B * T} vectors of one step per CPU step; oracle/vq_oracle_torch.py "
                      f"(the reference's ATen op sequence: N x K fp32 distances, one-hot, 3 sgemm), best of torch threads in {{all host cores, 32}}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the reference arm is entitled to every host thread
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(var, None)
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    v, ms = time_cpu(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "vectors/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "cpu_sample_vectors": CPU_SAMPLE_VECTORS},
        "cpu_baseline": cpu_baseline_block(v),
        "e2e": {"value": v, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------

def ncu_dram_bytes():
    """dram read + write bytes of one vq_assign_kernel launch, from the committed ncu summary (None if it is missing)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_assign_final_summary.txt")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    try:
        for line in open(path):
            if "dram__bytes_read.sum =" in line or "dram__bytes_write.sum =" in line:
                val, unit = line.split("=")[1].split()[:2]
                total += float(val) * mult[unit]
                seen += 1
    except Exception:
        return None
    return total if seen == 2 else None


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region, in-process through NVML.

    A forked `nvidia-smi -lms` loop was measured to perturb the very region it watches: one query can hold a driver lock
    for tens of milliseconds, and a 20-step timed region is only ~7 ms long (observed: 0.34 -> 1.5 / 3.1 ms per step when
    a query landed inside it).  NVML calls from a thread cost microseconds and fork nothing.
    """
    PERIOD_S = 0.002

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []          # (time, sm_mhz, reasons bitmask)
        self.smax = None
        self.handle = None
        self.stop_flag = False
        self.thread = None
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes torch's visible devices; honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis:
                parts = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(parts) and parts[self.gpu].isdigit():
                    idx = int(parts[self.gpu])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((time.time(), sm, rs))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def stop(self, t0, t1):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        if not inside:  # the timed region was shorter than one sample period: take the nearest samples
            inside = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
        sm = [r[1] for r in inside]
        reasons = sorted({n for r in inside for n, bit in names if r[2] & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                "samples": len(sm), "source": "nvml, in-process, 2 ms period"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import vector_quantize_pytorch_b200 as vqb
    from vector_quantize_pytorch_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(1234)  # same codebook on every rank (replicas)
    vq = vqb.VectorQuantize(dim=D, codebook_size=K, sync_codebook=world > 1).to(dev)
    with torch.no_grad():
        e = torch.randn(1, K, D, device=dev)
        vq._codebook.embed.copy_(e)
        vq._codebook.embed_avg.copy_(e)
    vq.train()
    gen = torch.Generator().manual_seed(1234 + rank)  # every rank its own shard of the global batch
    x_host = torch.randn(B, T, D, generator=gen).bfloat16().pin_memory()
    x_dev = x_host.to(dev)
    n_vec = B * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------- device-resident timing (`value`) with per-kernel events for the roofline
    # W untimed warm-up steps as requested, plus enough extra untimed calls for the allocator / graph cache to reach
    # their steady state (every output-pointer set is enqueued directly once and captured once before it replays)
    for _ in range(max(args.warmup, 12)):
        q, ind, loss = vq(x_dev)   # same binding pattern as the timed loop: the allocator then cycles the same blocks
    barrier()
    ops.PROFILE_EVENTS = None
    ops.LAUNCHES = 0
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.time()
    e0.record()
    for _ in range(args.steps):
        q, ind, loss = vq(x_dev)
    t_host = time.time()
    e1.record()
    barrier()
    t_end = time.time()
    launches = ops.LAUNCHES
    clocks = sampler.stop(t_start, t_end)
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    host_ms = (t_host - t_start) * 1e3 / args.steps   # CPU time to enqueue one step (must stay below ms_dev)

    # ---------------- the dominant kernel's launch duration (roofline): the same K steps once more, now with a CUDA
    # event pair recorded on the launching stream around vq_assign_kernel.  Kept out of the headline region because
    # event records cannot live inside the step's CUDA graph: with them every launch of the chain is enqueued one by
    # one and the step becomes sensitive to host jitter (observed 0.34 -> 1.1 ms on a noisy box).
    ops.PROFILE_EVENTS = []
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record()
    for _ in range(args.steps):
        q, ind, loss = vq(x_dev)
    g1.record()
    barrier()
    prof = ops.PROFILE_EVENTS
    ops.PROFILE_EVENTS = None
    ms_dev_events = g0.elapsed_time(g1) / args.steps
    prof = [pr for pr in prof if pr is not None]
    assign_ms = statistics.mean(a.elapsed_time(b) for a, b in prof) if prof else None

    # ---------------- end-to-end timing (`e2e`): pinned host input -> module -> host outputs
    q_host = torch.empty((B, T, D), dtype=torch.bfloat16).pin_memory()
    i_host = torch.empty((B, T), dtype=torch.int64).pin_memory()
    l_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        # public host-buffer API: pinned input -> chunk-pipelined H2D / kernels / D2H -> pinned outputs
        vq.forward_host(x_host, n_chunks=E2E_CHUNKS, out=(q_host, i_host, l_host))

    if os.environ.get("VQB_BENCH_SKIP_E2E"):  # profiling aid: keep the launch list to the device-resident steps
        print(json.dumps({"ms_per_step": ms_dev, "kernel_ms": assign_ms, "gpu_launches": launches, "host_ms_per_step": host_ms}))
        return
    for _ in range(max(3, min(args.warmup, 5))):   # >= 3: every chunk's pointer set is seen twice before it replays
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1) / args.steps)
    h2d = x_host.numel() * x_host.element_size()
    d2h = q_host.numel() * 2 + i_host.numel() * 8 + 4

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    flops = 2.0 * n_vec * K * D  # algorithmic: one pass of the N x K x D contraction
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    roof = {"bound": "tensor", "kernel": "vq_assign_kernel (tcgen05 distance MMA + fused arg-max)",
            "achieved": flops / (assign_ms * 1e-3) / 1e12 if assign_ms else None, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (flops / (assign_ms * 1e-3) / 1e12 / peak_tf) if assign_ms else None,
            "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside the step)",
            "kernel_ms": assign_ms, "kernel_share_of_step": assign_ms / ms_dev if assign_ms else None,
            "measured": "CUDA event pair on the launching stream around every vq_assign_kernel launch, over the same K "
                        "steps repeated right after the headline region (events split the step's CUDA graph)",
            "ms_per_step_with_events": ms_dev_events,
            "algorithmic_flops_per_launch": flops, "executed_mma_passes": 2, "traffic": ncu_dram_bytes(),
            "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_assign_final_summary.txt: "
                            "`ncu --set full` capture of the search launch WITHOUT the fused tail; in the step the same launch "
                            "also writes quantize, +134.2 MB)"}
    cpu_v, _ = time_cpu(steps=2, warmup=1)
    line = {
        "metric": METRIC, "value": world * n_vec / (ms_dev * 1e-3), "unit": "vectors/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_vectors": n_vec, "global_vectors": world * n_vec,
                   "parallelism": f"dp{world}: batch sharded, one NCCL all-reduce of packed EMA stats per step" if world > 1 else "single GPU",
                   "l2": "input (134 MB) + output (134 MB) per step exceed the 126 MB L2; no extra flush",
                   "index_mismatch_policy": "bit-exact vs oracle outside fp32 near-ties (tests/test_parity_gpu.py)"},
        "e2e": {"value": world * n_vec / (ms_e2e * 1e-3), "unit": "vectors/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "host_ms_per_step": host_ms,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu_baseline_block(cpu_v),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
