"""bench.py -- the driver's measurement contract for the DeepJ hot path.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Workload (BASELINE.json configs[2], "DeepJ training, batch 64 synthetic
sequences, data-parallel with NCCL gradient allreduce"): one step = forward +
primary_loss + backward + gradient exchange + Nadam (one fused kernel over NVLink
peer memory; DJ_PEER_NADAM=0 = NCCL all-reduce + Nadam kernel) on 64 synthetic
[128,48,3] windows PER GPU (weak scaling), default constants.py model, dropout
on, bf16 gate-GEMM operands / fp32 everything else.  A short generation probe
(configs[1], 1 style-mixed sequence) is reported under "generation".

The reference's Keras/TensorFlow path cannot run here (not installable, see
DESIGN.md); `--impl reference` and `cpu_baseline` time the CPU oracle -- a
restatement of model.py -- on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "train_seqs_per_sec", "seqs/s"
BATCH = 64          # per GPU
REF_SAMPLE_B = 2    # sequences per CPU step (bounded sample of the same workload)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region: NVML every 10 ms (an nvidia-smi process per
    sample is too slow for a 100 ms region), nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        self.rows.append([sm, self.max_sm] + ["Active" if (r & b) else "Not Active" for b in bits])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 6:
            self.rows.append(f)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(r[2 + i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_oracle_rate(steps, warmup, B):
    """Times the CPU oracle's train step (forward + loss + autograd backward +
    Nadam) on `B` sequences per step; returns (seqs/s of the median step, threads)."""
    from oracle import deepj_oracle as O
    cfg = O.Config()
    p = O.init_params(cfg, 0)
    batch = O.synthetic_batch(cfg, B, cfg.seq_len, 1234)
    masks = O.random_masks(cfg, B, cfg.seq_len, 7)
    st = O.NadamState()
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        _, _, grads = O.loss_and_grads(p, cfg, *batch, masks)
        p = O.nadam_step(p, grads, st)
        if i >= warmup:
            times.append(time.time() - t0)
    times.sort()
    return B / times[len(times) // 2], torch.get_num_threads()      # median step


def workload_config(scaled: bool, B: int, T: int, world: int) -> dict:
    """`config` of the JSON line: the same for our arm and for the reference arm."""
    return {"workload": ("Scaled biaxial LSTM (BASELINE configs[4]): 512/256 units, 512-step windows, "
                         "synthetic batch per GPU, fwd+loss+bwd+gradient exchange+Nadam, dropout on") if scaled else
                        ("DeepJ training (BASELINE configs[2]): default constants.py model, "
                         "batch 64 synthetic sequences per GPU, fwd+loss+bwd+gradient exchange+Nadam, dropout on"),
            "batch_per_gpu": B, "global_batch": B * world, "seq_len": T, "parallelism": f"dp{world}",
            "l2": "per-step working set (>5 GB of activations) exceeds the 126 MB L2; no explicit flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1)
    steps = max(1, min(args.steps, 6))       # each CPU step is seconds; keep the arm within minutes
    rate, threads = cpu_oracle_rate(steps, warm, REF_SAMPLE_B)
    sample = (f"{steps} CPU steps of {REF_SAMPLE_B} sequences (forward+loss+backward+Nadam, fp32 torch-CPU oracle; "
              f"Keras/TF reference not installable)")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * REF_SAMPLE_B / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same workload as our arm; every CPU step is a bounded sample of it (REF_SAMPLE_B of the 64 sequences)
            "config": dict({k: v for k, v in workload_config(False, BATCH, 128, 1).items() if k != "l2"},
                           sample_sequences_per_step=REF_SAMPLE_B),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class StdoutToStderr:
    """NCCL prints its version banner on the C-level stdout at first use; the contract is ONE JSON line
    on stdout, so everything before the final print goes to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def run_ours(args):
    with StdoutToStderr():
        line = _run_ours(args)
    if line is not None:
        print(json.dumps(line), flush=True)


def _run_ours(args):
    import torch.distributed as dist
    import music_generator_b200  # noqa: F401
    from music_generator_b200.config import ModelConfig
    from music_generator_b200.engine import Engine
    from music_generator_b200.sampler import generate_events
    import dataset

    from music_generator_b200 import parallel
    rank, world, local = parallel.env_world()
    torch.cuda.set_device(local)
    parallel.init_distributed("nccl")
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    scaled = args.workload == "scaled"
    T = 512 if scaled else 128
    mcfg = ModelConfig(time_axis_units=512, note_axis_units=256, seq_len=512) if scaled else ModelConfig()
    if scaled and args.batch == BATCH:
        B = 16                                   # BASELINE configs[4]: 2x hidden units, 4x sequence length, B=16/GPU
    eng = Engine(mcfg, precision="bf16")
    eng.init_params(0)
    x, y = dataset.synthetic_all(B, T, seed=1234 + rank)
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x[0], x[1], x[2], x[3], y[0])]
    dev = [h.cuda(non_blocking=True) for h in host]
    # gradient exchange: NCCL all-reduce + Nadam kernel, or (DJ_PEER_NADAM=1) one fused kernel over peer memory
    allreduce, peer = parallel.make_step_exchange(eng, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return parallel.max_over_ranks(ms, "cuda")

    # ---------------- device-resident arm (value)
    for i in range(W):
        eng.train_step(*dev, seed=i, allreduce=allreduce, world=world)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    DOM = "dj_lstm_scan_tc_bwd:bwd:time1"
    eng.profile, eng.profile_only = [], {DOM}
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = eng.train_step(*dev, seed=100 + i, allreduce=allreduce, world=world)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launches - launches0
    dom = eng.profile_summary().get(DOM, [0, 0.0])
    eng.profile, eng.profile_only = None, None
    value = world * B * K / (ms * 1e-3)

    # ---------------- end-to-end arm: host buffers in, loss out, every step
    # Every step's inputs come from pinned host memory and every step's loss is read back (a sync).  The copy of
    # step i+1's batch is issued on a copy stream while step i computes, as the fit() input pipeline does.
    h2d = sum(h.numel() * h.element_size() for h in host)
    copy_stream = torch.cuda.Stream()
    bufs = [[torch.empty_like(t) for t in dev] for _ in range(2)]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(bufs[slot], host):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_steps(n, seed0):
        ev = prefetch(0)
        last = 0.0
        for i in range(n):
            torch.cuda.current_stream().wait_event(ev)
            d = bufs[i & 1]
            loss_t = eng.train_step(*d, seed=seed0 + i, allreduce=allreduce, world=world)
            if i + 1 < n:
                ev = prefetch((i + 1) & 1)      # that slot fed step i-1, whose loss has been read back: it is free
            last = float(loss_t.item())
        return last

    e2e_steps(2, 0)
    barrier()
    e0.record()
    lossv = e2e_steps(K, 200)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e = world * B * K / (ms_e2e * 1e-3)
    exchange = "NCCL all-reduce of the flat gradient + Nadam kernel"
    if peer is not None:
        exchange = "fused reduce-scatter + Nadam + all-gather kernel over peer memory (CUDA IPC)"
        peer.raise_if_timed_out()
        peer.close(eng)          # collective: every rank, before the non-zero ranks leave

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    # ---------------- roofline of the dominant kernel (time-axis reverse scan, layer 1)
    # algorithmic bytes per row: gates 16U + c 4U + dY 4U read, dZ(bf16) 8U written, U=256
    U, M = mcfg.time_axis_units, B * T * 48
    alg_bytes = 32 * U * M
    peak, how = peaks()
    dom_ms = dom[1] / max(dom[0], 1)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
    kname = "scan_tc_bwd_kernel<512,16,32,time>" if scaled else "scan_tc_bwd_kernel<256,48,64,time>"
    roofline = {"kernel": kname + " (dj_lstm_scan_tc_bwd, time-axis layer 1 reverse scan)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the ncu --set full
                # capture summarised in profiles/r01_scan_tc_bwd_time.md (2.418 GB + 0.786 GB per launch)
                "traffic": 3.204e9 if (B == 64 and not scaled) else None, "peak_source": how, "avg_launch_ms": dom_ms, "launches_timed": dom[0],
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "sequential recurrence: bound by the per-step publish/TMA/MMA latency chain, not by HBM; timed inside the "
                        "step, where it shares the GPU with the weight-gradient GEMMs of the second stream (alone: 0.87 ms), "
                        "see DESIGN.md section 4"}

    # ---------------- generation probe (configs[1])
    gen = None
    if not args.no_generation and not scaled:
        ge = Engine(ModelConfig(), precision="fp32")
        ge.init_params(0)
        sty = np.mean([np.eye(23)[i] for i in (0, 5, 12)], axis=0)
        gsteps = 32
        u = np.random.RandomState(42).random_sample(2 * 48 * gsteps)
        generate_events(ge, [sty], 4, u)
        torch.cuda.synchronize()
        rates = []
        for _ in range(3):                       # the one-sequence loop is launch-heavy: median of three runs
            t0 = time.time()
            generate_events(ge, [sty], gsteps, u)
            torch.cuda.synchronize()
            rates.append(gsteps / (time.time() - t0))
        gen = {"timesteps_per_s": sorted(rates)[1], "sequences": 1, "timesteps": gsteps, "runs": [round(r, 1) for r in rates],
               "workload": "generate.py path, 1 style-mixed sequence, full 128-step window recompute per timestep"}
        # configs[3] per-GPU share: 128 independent sequences (4 predict-chunks of 32), indexed uniform stream
        Gb, bsteps = 128, 4
        stys = [np.eye(23)[i % 23] for i in range(Gb)]
        ub = np.random.RandomState(7).random_sample((bsteps, Gb, 48, 2))
        generate_events(ge, stys, 1, ub[:1], stream_mode=1)
        torch.cuda.synchronize()
        t0 = time.time()
        generate_events(ge, stys, bsteps, ub, stream_mode=1)
        torch.cuda.synchronize()
        gen["batched"] = {"sequences": Gb, "timesteps": bsteps, "timesteps_per_s": Gb * bsteps / (time.time() - t0)}

    # ---------------- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, threads = cpu_oracle_rate(5, 1, REF_SAMPLE_B)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"median of 5 steps of {REF_SAMPLE_B} sequences after 1 warm-up, fp32 torch-CPU oracle of model.py"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 gate-GEMM operands, f32 accumulate/recurrence", "data": "synthetic",
            "config": workload_config(scaled, B, T, world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "generation": gen, "loss": lossv}
    if world > 1:
        line["config"]["exchange"] = exchange
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--workload", default="default", choices=["default", "scaled"])
    ap.add_argument("--no-generation", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
