"""bench.py -- the driver's measurement contract for the DeepJ hot path.

  python bench.py --gpus N --steps K --warmup W                     (our arm; torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W    (CPU reference arm, rank 0 only)

Workloads (`--workload`, BASELINE.json configs):
  train   (default, configs[2]) one step = forward + primary_loss + backward + gradient exchange + Nadam on 64
          synthetic [128,48,3] windows PER GPU (`--scaling weak`, default) or 64 in total (`--scaling strong`),
          default constants.py model, dropout on.  `--batch 16` is configs[0]/[1]'s batch size.
  scaled  (configs[4]) the same step on the 2x-units / 4x-length model, 16 windows per GPU.
  gen1    (configs[1]) autoregressive generation of ONE style-mixed sequence, 512 timesteps, reference uniform
          stream; a "step" is one generated timestep.
  gen1024 (configs[3]) 128 independent sequences PER GPU (1024 on 8), indexed uniform stream, no collective.

The default line also carries a `generation` object (gen1 + a 128-sequence probe + the CPU oracle in the
reference's literal call structure) so that one driver run reports both halves of BASELINE's metric.

The reference's Keras/TensorFlow path cannot run here (not installable); `--impl reference` and
`cpu_baseline` time the CPU oracle -- a restatement of model.py / generate.py -- on the host cores
(`kind: "port"`), with every host thread, whatever OMP_NUM_THREADS torchrun exported.
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

BATCH = 64          # sequences per GPU (weak scaling) / in total (strong scaling)
REF_SAMPLE_B = 16   # sequences per CPU step: the reference's own batch size (constants.py:66)
N_NOTES = 48


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def measured_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures summarised under
    profiles/ (profiles/traffic.json: kernel/shape key -> bytes, with the capture it came from), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    ent = json.load(open(p)).get(kernel_key)
    return (ent["dram_bytes_per_launch"], ent["source"]) if ent else (None, None)


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

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.summary()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(r[2 + i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ----------------------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is one process and takes every core the box has."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def oracle_config(scaled):
    from oracle import deepj_oracle as O
    return O.Config(time_axis_units=512, note_axis_units=256, seq_len=512) if scaled else O.Config()


def cpu_train_rate(steps, warmup, B, scaled=False):
    """CPU oracle train step (forward + loss + autograd backward + Nadam) on `B` sequences per step;
    returns (seqs/s of the median step, threads)."""
    from oracle import deepj_oracle as O
    threads = use_all_host_threads()
    cfg = oracle_config(scaled)
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
    return B / times[len(times) // 2], threads


def cpu_generation_rate(G, timesteps, warmup=1):
    """CPU oracle generation in the reference's LITERAL call structure (generate.py:104-121: one full-window
    time_model.predict and 48 full note_model.predict per timestep) for G sequences; returns (timesteps/s, threads)."""
    from oracle import deepj_oracle as O
    threads = use_all_host_threads()
    cfg = O.Config()
    p = O.init_params(cfg, 0)
    styles = [np.mean([np.eye(23)[i] for i in (0, 5, 12)], axis=0)] if G == 1 else \
        [np.eye(23)[i % 23] for i in range(G)]
    u = np.random.RandomState(42).random_sample(2 * N_NOTES * (timesteps + warmup) * G)
    O.generate(p, cfg, styles, warmup, u, mode="literal")
    t0 = time.time()
    O.generate(p, cfg, styles, timesteps, u, mode="literal")
    return G * timesteps / (time.time() - t0), threads


def workload_config(args, B, T, world):
    """`config` of the JSON line: the same for our arm and for the reference arm."""
    names = {
        "train": "DeepJ training (BASELINE configs[2]): default constants.py model, synthetic sequences, "
                 "fwd+loss+bwd+gradient exchange+Nadam, dropout on",
        "scaled": "Scaled biaxial LSTM (BASELINE configs[4]): 512/256 units, 512-step windows, synthetic batch per GPU, "
                  "fwd+loss+bwd+gradient exchange+Nadam, dropout on",
        "gen1": "DeepJ generation (BASELINE configs[1]): 1 style-mixed sequence, 512 timesteps, reference uniform stream, "
                "full 128-step window recompute per timestep",
        "gen1024": "DeepJ batched generation (BASELINE configs[3]): 128 independent style-conditioned sequences per GPU "
                   "(1024 on 8), indexed uniform stream, no collective",
    }
    cfg = {"workload": names[args.workload], "parallelism": f"dp{world}"}
    if args.workload in ("train", "scaled"):
        cfg.update(batch_per_gpu=B, global_batch=B * world, seq_len=T,
                   l2="per-step working set (>5 GB of activations) exceeds the 126 MB L2; no explicit flush")
    else:
        cfg.update(sequences_per_gpu=B, sequences=B * world, window=128,
                   l2="every timestep rewrites the 100+ MB window activations per 32 sequences; no explicit flush")
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    warm = min(args.warmup, 1)
    if args.workload in ("train", "scaled"):
        scaled = args.workload == "scaled"
        sample_b = 2 if scaled else REF_SAMPLE_B
        steps = max(1, min(args.steps, 3 if scaled else 5))     # each CPU step is seconds; keep the arm within minutes
        rate, threads = cpu_train_rate(steps, warm, sample_b, scaled)
        B, T = (16, 512) if scaled else (BATCH if args.scaling == "weak" else BATCH // world, 128)
        metric, unit, ms = "train_seqs_per_sec", "seqs/s", 1e3 * sample_b / rate
        sample = (f"median of {steps} CPU steps of {sample_b} sequences each (forward+loss+autograd backward+Nadam, fp32 "
                  f"torch-CPU oracle of model.py; Keras/TF reference not installable), {threads} threads")
    else:
        G = 1 if args.workload == "gen1" else 4
        steps = max(2, min(args.steps, 8))
        rate, threads = cpu_generation_rate(G, steps)
        B, T = (1 if args.workload == "gen1" else 128), 128
        metric, unit, ms = "generated_timesteps_per_sec", "timesteps/s", 1e3 * G / rate
        sample = (f"{steps} timesteps of {G} sequence(s), fp32 torch-CPU oracle in the reference's literal call structure "
                  f"(1 time_model.predict + 48 note_model.predict per timestep, generate.py:104-121), {threads} threads")
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B, T, world),
            "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
        line = _run_train(args) if args.workload in ("train", "scaled") else _run_generation(args)
    if line is not None:
        print(json.dumps(line), flush=True)


def _dist_helpers(world):
    import torch.distributed as dist
    from music_generator_b200 import parallel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(ms):
        """(max, min, all) of a per-rank device time."""
        if world <= 1:
            return ms, ms, [ms]
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        v = [float(x.item()) for x in out]
        return max(v), min(v), v
    return barrier, over_ranks, parallel


# ----------------------------------------------------------------------------------------------------------------
# algorithmic HBM bytes of one training step in the layout the kernels actually use
# ----------------------------------------------------------------------------------------------------------------
def step_algorithmic_bytes(mcfg, B, T, mixed):
    """Per LSTM layer and activation row (K = padded input width, U units, Up = units of the layer below):
      forward   producer reads h_below 4Up, writes A (bf16 hi [+ lo]) 2K [4K]; gate GEMM reads it back 2K [4K], writes
                Z fp32 16U; scan reads Z 16U, writes gates (4 halves) 8U + h 4U + c 4U + 16-bit h_{t-1} 2U
      backward  scan reads gates 8U + c 4U + dY 4U, writes bf16 dZ 8U; data-gradient GEMM reads dZ 8U, writes dA fp32 4K;
                weight-gradient GEMMs read A_hi 2K + dZ 8U and h_{t-1} 2U + dZ 8U; style reduction reads dA 4K
      = (14K [18K]) + 100U + 4Up bytes per row; heads add 4U read + 4U dX written + 24 B; rows = B*T*48."""
    rows = B * T * N_NOTES
    total = 0
    up = 0
    for L in mcfg.layers():
        K = (L["F"] + 31) // 32 * 32
        U = L["U"]
        total += (18 if mixed else 14) * K + 100 * U + 4 * up
        up = U
    total += 8 * mcfg.note_axis_units + 24
    return total * rows


def _run_train(args):
    import torch.distributed as dist
    import music_generator_b200  # noqa: F401
    from music_generator_b200.config import ModelConfig
    from music_generator_b200.engine import Engine
    import dataset

    from music_generator_b200 import parallel
    rank, world, local = parallel.env_world()
    torch.cuda.set_device(local)
    parallel.init_distributed("nccl")
    barrier, over_ranks, _ = _dist_helpers(world)
    K, W = args.steps, max(args.warmup, 3)
    scaled = args.workload == "scaled"
    T = 512 if scaled else 128
    mcfg = ModelConfig(time_axis_units=512, note_axis_units=256, seq_len=512) if scaled else ModelConfig()
    B = args.batch if args.batch else (16 if scaled else BATCH)
    if args.scaling == "strong":
        assert B % world == 0, "strong scaling splits the global batch evenly"
        B //= world
    eng = Engine(mcfg, precision=args.precision)
    eng.init_params(0)
    x, y = dataset.synthetic_all(B, T, seed=1234 + rank)
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x[0], x[1], x[2], x[3], y[0])]
    dev = [h.cuda(non_blocking=True) for h in host]
    # gradient exchange: one fused kernel over peer memory, or (DJ_PEER_NADAM=0) NCCL all-reduce + Nadam kernel
    allreduce, peer = parallel.make_step_exchange(eng, world)

    # ---------------- device-resident arm (value)
    for i in range(W):
        eng.train_step(*dev, seed=i, allreduce=allreduce, world=world)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for i in range(K):
        loss = eng.train_step(*dev, seed=100 + i, allreduce=allreduce, world=world)
    host_ms = (time.perf_counter() - t_host) * 1e3 / K     # host time to ENQUEUE a step (no sync inside the loop)
    e1.record()
    barrier()
    ms, ms_min, ms_all = over_ranks(e0.elapsed_time(e1))
    launches = eng.launches - launches0
    value = world * B * K / (ms * 1e-3)
    graphed = bool(eng._graphs) and all(g["graph"] is not None for g in eng._graphs.values())

    # ---------------- the dominant kernel, timed inside the step with CUDA events on its launching stream.  A graph
    # replay has no place for host-recorded events between its nodes, so these K steps are launched from Python (same
    # kernels, same two-stream backward); their step time is reported beside the graph's.
    DOM = "dj_lstm_scan_tc_bwd:bwd:time1"
    eng.profile, eng.profile_only = [], {DOM}
    e0.record()
    for i in range(K):
        eng.train_step(*dev, seed=150 + i, allreduce=allreduce, world=world)
    e1.record()
    barrier()
    ms_eager, _, _ = over_ranks(e0.elapsed_time(e1))
    dom = eng.profile_summary().get(DOM, [0, 0.0])
    eng.profile, eng.profile_only = None, None

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

    # Every step's loss is read on the host inside the timed region, the way a training loop logs it: copied to
    # pinned memory behind the step and read ONE step later, while the next step is already running (reading it before
    # enqueueing the next step would idle the GPU for the host's enqueue time every step).  The last one is read at the end.
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_steps(n, seed0):
        ev = prefetch(0)
        last = 0.0
        for i in range(n):
            torch.cuda.current_stream().wait_event(ev)
            d = bufs[i & 1]
            loss_t = eng.train_step(*d, seed=seed0 + i, allreduce=allreduce, world=world)
            loss_host[i & 1:(i & 1) + 1].copy_(loss_t.reshape(1), non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()          # step i-1 is complete: its batch slot is free
                last = float(loss_host[(i - 1) & 1])
            if i + 1 < n:
                ev = prefetch((i + 1) & 1)      # that slot fed step i-1, which has completed
        loss_ev[(n - 1) & 1].synchronize()
        return float(loss_host[(n - 1) & 1])

    e2e_steps(2, 0)
    barrier()
    e0.record()
    lossv = e2e_steps(K, 200)
    e1.record()
    barrier()
    ms_e2e, _, _ = over_ranks(e0.elapsed_time(e1))
    clocks = sampler.finish()
    e2e = world * B * K / (ms_e2e * 1e-3)

    # ---------------- per-entry-point device times of 3 more steps (events around every launch: the two backward
    # streams are serialised here, so these are kernel times, not the step's critical path)
    kernels = None
    if rank == 0 and not args.no_kernel_table:
        eng.profile, eng.profile_only = [], None
    if not args.no_kernel_table:
        for i in range(3):
            eng.train_step(*dev, seed=300 + i, allreduce=allreduce, world=world)
        barrier()
    if rank == 0 and not args.no_kernel_table:
        agg = eng.profile_summary()
        eng.profile = None
        kernels = {k: {"launches_per_step": v[0] / 3, "ms_per_step": round(v[1] / 3, 4)}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]}

    exchange = "NCCL all-reduce of the flat gradient + Nadam kernel"
    if peer is not None:
        exchange = "fused reduce-scatter + Nadam + all-gather kernel over peer memory (CUDA IPC)"
        peer.raise_if_timed_out()
        peer.close(eng)          # collective: every rank, before the non-zero ranks leave

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    # ---------------- roofline of the dominant kernel (time-axis reverse scan, layer 1) and of the whole step
    # algorithmic bytes per row of that kernel: gates (4 halves) 8U + c 4U + dY 4U read, dZ (bf16) 8U written
    U, M = mcfg.time_axis_units, B * T * N_NOTES
    alg_bytes = 24 * U * M
    peak, how = peaks()
    dom_ms = dom[1] / max(dom[0], 1)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
    kname = f"scan_tc_bwd_kernel<{U},time> B={B} T={T}"
    traffic, traffic_src = measured_traffic(kname)
    roofline = {"kernel": kname + " (dj_lstm_scan_tc_bwd, time-axis layer 1 reverse scan)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": how, "avg_launch_ms": dom_ms,
                "launches_timed": dom[0], "algorithmic_bytes_per_launch": alg_bytes,
                "note": "sequential recurrence: every step each SM must ingest the multicast all-gather of the tile's dz "
                        "(2 KB per sequence and step) before its MMAs can run; that on-chip exchange and the per-step "
                        "fence/TMA/MMA latency chain bound the kernel, not HBM (DESIGN.md section 4, "
                        "profiles/r02_scan_bwd_trace.md); timed inside the step, where it shares the GPU with the "
                        "weight-gradient GEMMs of the second stream"}
    step_bytes = step_algorithmic_bytes(mcfg, B, T, args.precision == "mixed")
    step_roofline = {"bound": "hbm", "algorithmic_bytes_per_step": step_bytes,
                     "achieved": step_bytes / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": step_bytes / (ms / K * 1e-3) / 1e9 / peak,
                     "formula": "(18K|14K)+100U+4Up bytes per row and LSTM layer (mixed|bf16), see step_algorithmic_bytes"}

    # ---------------- generation (configs[1], and one GPU's share of configs[3])
    gen = None
    if not args.no_generation and not scaled and world == 1:
        gen = generation_probe(args)

    # ---------------- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sb = 2 if scaled else REF_SAMPLE_B
        rate, threads = cpu_train_rate(3, 1, sb, scaled)
        cpu = {"value": rate, "unit": "seqs/s", "cores": threads, "kind": "port",
               "sample": f"median of 3 steps of {sb} sequences after 1 warm-up, fp32 torch-CPU oracle of model.py "
                         f"(forward+loss+autograd backward+Nadam), {threads} threads"}

    dtype = {"mixed": "bf16 hi+lo split gate-GEMM operands (fp32-grade product), f16 h / f16 hi+lo U in the recurrence, "
                      "bf16 dZ in backward; f32 accumulate, state and activations",
             "bf16": "bf16 operands in the gate GEMMs AND the recurrence, f32 accumulate/state (fast mode, outside the 1e-3 "
                     "output tolerance)",
             "fp32": "f32"}[args.precision]
    line = {"metric": "train_seqs_per_sec", "value": value, "unit": "seqs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": dtype, "precision": args.precision, "data": "synthetic",
            "config": workload_config(args, B, T, world),
            "e2e": {"value": e2e, "unit": "seqs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K,
                    "how": "per step: that step's batch copied from pinned host memory on a copy stream (double-buffered), "
                           "the step, its loss copied to pinned host memory and read by the host one step later"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "step_roofline": step_roofline,
            "rank_ms_per_step": {"max": ms / K, "min": ms_min / K, "all": [round(v / K, 4) for v in ms_all]},
            # host time to enqueue one step of the timed region (one graph launch + the 128-byte parameter upload when
            # cuda_graph is true, ~56 kernel launches from Python otherwise)
            "host_enqueue_ms_per_step": host_ms, "launches_per_step": launches / K,
            "cuda_graph": graphed, "ms_per_step_launched_from_python": ms_eager / K,
            "kernels": kernels, "cpu_baseline": cpu, "generation": gen, "loss": lossv}
    if world > 1:
        line["config"]["exchange"] = exchange
        dist.destroy_process_group()
    return line


# ----------------------------------------------------------------------------------------------------------------
# generation workloads
# ----------------------------------------------------------------------------------------------------------------
def _gen_styles(G, seed=0):
    """the three genre mixtures (dataset.compute_genre) cycled with random 3-composer mixtures."""
    import dataset
    rs = np.random.RandomState(seed)
    out = []
    for g in range(G):
        if g % 2 == 0:
            out.append(np.asarray(dataset.compute_genre((g // 2) % 3), dtype=np.float64))
        else:
            out.append(np.mean([np.eye(23)[i] for i in rs.choice(23, 3, replace=False)], axis=0))
    return out


def _timed_generation(eng, styles, steps, warmup, u, stream_mode):
    """Device time of `steps` generated timesteps after `warmup` untimed ones (state carried over), CUDA events."""
    from music_generator_b200.sampler import GenerationRun
    run = GenerationRun(eng, styles, warmup + steps, u, stream_mode)
    for t in range(warmup):
        run.step(t)
    torch.cuda.synchronize()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(warmup, warmup + steps):
        run.step(t)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), eng.launches - l0, run


def generation_probe(args):
    """configs[1] (one sequence, 512 timesteps) and one GPU's share of configs[3] (128 sequences), with the CPU oracle's
    literal loop beside them and a lock-step bit-exactness check of the first timesteps."""
    from music_generator_b200.config import ModelConfig
    from music_generator_b200.engine import Engine
    from music_generator_b200.sampler import generate_events
    from oracle import deepj_oracle as O
    ge = Engine(ModelConfig(), precision="fp32")
    ge.init_params(0)
    sty = np.mean([np.eye(23)[i] for i in (0, 5, 12)], axis=0)
    gsteps, check = 512, 16
    u = np.random.RandomState(42).random_sample(2 * N_NOTES * (gsteps + 4))
    rates = []
    for _ in range(3):                       # the one-sequence loop is launch-heavy: median of three runs
        ms, launches, run = _timed_generation(ge, [sty], gsteps, 4, u, 0)
        rates.append(gsteps / (ms * 1e-3))
    # bit-exactness: the oracle replays the first `check` timesteps in lock-step on the device's events
    ev, info = generate_events(ge, [sty], check, u, stream_mode=0)
    p32 = {k: torch.tensor(v) for k, v in ge.get_params().items()}
    _, oinfo = O.generate(p32, O.Config(), [sty], check, u, mode="incremental", forced_events=ev, return_probs=True)
    exact = bool(np.array_equal(oinfo["decisions"][..., :2], ev[..., :2]) and info["uniforms_used"] == oinfo["uniforms_used"])
    gen = {"timesteps_per_s": sorted(rates)[1], "sequences": 1, "timesteps": gsteps, "runs": [round(r, 1) for r in rates],
           "gpu_launches_per_timestep": launches / gsteps,
           "events_equal_oracle": exact, "events_checked_timesteps": check,
           "max_prob_err_vs_oracle": float(np.abs(info["probs"] - oinfo["probs"]).max()),
           "workload": "generate.py path, 1 style-mixed sequence, 32 bars, full 128-step window recompute per timestep"}
    Gb, bsteps = 128, 6
    ub = np.random.RandomState(7).random_sample((bsteps + 2, Gb, N_NOTES, 2))
    ms, launches, _ = _timed_generation(ge, _gen_styles(Gb), bsteps, 2, ub, 1)
    gen["batched"] = {"sequences": Gb, "timesteps": bsteps, "timesteps_per_s": Gb * bsteps / (ms * 1e-3),
                      "gpu_launches_per_timestep": launches / bsteps}
    if not args.no_cpu_baseline:
        rate, threads = cpu_generation_rate(1, 6)
        gen["cpu_baseline"] = {"value": rate, "unit": "timesteps/s", "cores": threads, "kind": "port",
                               "sample": "6 timesteps of 1 sequence after 1 warm-up, fp32 torch-CPU oracle in the reference's "
                                         "literal call structure (49 predicts per timestep, generate.py:104-121)"}
    return gen


def _run_generation(args):
    import torch.distributed as dist
    import music_generator_b200  # noqa: F401
    from music_generator_b200.config import ModelConfig
    from music_generator_b200.engine import Engine
    from music_generator_b200.sampler import generate_events
    from music_generator_b200 import parallel
    rank, world, local = parallel.env_world()
    torch.cuda.set_device(local)
    parallel.init_distributed("nccl")        # only for the barrier and the max over ranks: generation has no collective
    barrier, over_ranks, _ = _dist_helpers(world)
    single = args.workload == "gen1"
    K = args.steps if args.steps_given else (512 if single else 32)
    W = max(args.warmup, 3)
    G = 1 if single else 128
    eng = Engine(ModelConfig(), precision="fp32")
    eng.init_params(0)
    if single:
        styles = [np.mean([np.eye(23)[i] for i in (0, 5, 12)], axis=0)]
        u = np.random.RandomState(42).random_sample(2 * N_NOTES * (K + W))
        mode = 0
    else:
        # every rank draws the WHOLE batch's indexed stream and keeps its own sequences (what generate_batch does)
        styles = _gen_styles(G * world)[rank * G:(rank + 1) * G]
        u = np.random.RandomState(7).random_sample((K + W, G * world, N_NOTES, 2))[:, rank * G:(rank + 1) * G]
        u = np.ascontiguousarray(u)
        mode = 1
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_local, launches, run = _timed_generation(eng, styles, K, W, u, mode)
    barrier()
    ms, ms_min, ms_all = over_ranks(ms_local)
    value = world * G * K / (ms * 1e-3)
    # ---- end to end through the public call: numpy styles + uniforms in, numpy events out (H2D / D2H inside)
    generate_events(eng, styles, 2, u[:2] if mode == 1 else u, stream_mode=mode)
    barrier()
    t0 = time.time()
    ev, _ = generate_events(eng, styles, K, u[:K] if mode == 1 else u, stream_mode=mode)
    torch.cuda.synchronize()
    ms_e2e, _, _ = over_ranks((time.time() - t0) * 1e3)
    clocks = sampler.finish()
    e2e = world * G * K / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        Gc = 1 if single else 4
        rate, threads = cpu_generation_rate(Gc, 6)
        cpu = {"value": rate, "unit": "timesteps/s", "cores": threads, "kind": "port",
               "sample": f"6 timesteps of {Gc} sequence(s) after 1 warm-up, fp32 torch-CPU oracle in the reference's literal "
                         f"call structure (generate.py:104-121), {threads} threads"}
    # algorithmic work of one timestep of one sequence: 10.93 GFLOP, sequential depth 256 + 96 LSTM steps
    flops = 10.93e9 * G * K / (ms * 1e-3)
    if eng.gen_tc:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
        roof = {"bound": "tensor", "peak": pk / 3.0,
                "note": "fp32-grade products on the tensor cores cost three half-precision MMA passes (hi.hi + lo.hi + hi.lo), "
                        "so the peak for ALGORITHMIC flops is the measured sustained 16-bit rate / 3; the sampled events must "
                        "equal the fp32 model's, which rules out single-pass 16-bit operands.  At 1 sequence the path is a "
                        "latency chain of 352 dependent LSTM steps per timestep, not a throughput problem"}
    else:
        roof = {"bound": "fma", "peak": 74.4,
                "note": "peak = fp32 FMA rate of the CUDA cores (148 SMs x 128 lanes x 2 x 1.965 GHz)"}
    line = {"metric": "generated_timesteps_per_sec", "value": value, "unit": "timesteps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (random-init weights, seeded uniform stream)",
            "config": workload_config(args, G, 128, world),
            "e2e": {"value": e2e, "unit": "timesteps/s", "h2d_bytes_per_step": int(u.nbytes // max(K + W, 1)) if mode == 1 else 16 * N_NOTES,
                    "d2h_bytes_per_step": G * N_NOTES * 3 * 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": dict({"kernel": ("time-axis window recompute (2 LSTM layers x 128 steps per generated timestep): "
                                         + ("dj_gate_gemm_16s + dj_lstm_scan_tc_gen2 (both layers in one launch, layer 1 one "
                                            "step behind layer 0: ~130 sequential steps)" if (single and eng.gen_fused) else
                                            "dj_gate_gemm_16s + dj_lstm_scan_tc_infer")) if eng.gen_tc else
                                        "time-axis window recompute: dj_gemm_simt + dj_lstm_scan_fwd",
                              "achieved": flops / 1e12, "unit": "TFLOP/s", "frac": flops / 1e12 / roof["peak"],
                              "traffic": None}, **roof),
            "rank_ms_per_step": {"max": ms / K, "min": ms_min / K, "all": [round(v / K, 4) for v in ms_all]},
            "cpu_baseline": cpu, "played_notes": int(ev[..., 0].sum())}
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU (default 64; 16 for the scaled model)")
    ap.add_argument("--workload", default="train", choices=["train", "default", "scaled", "gen1", "gen1024"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="mixed", choices=["mixed", "bf16"])
    ap.add_argument("--no-generation", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
    args = ap.parse_args()
    if args.workload == "default":
        args.workload = "train"
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 10
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
