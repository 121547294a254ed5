"""Data-parallel plumbing (one process per GPU): rendezvous, batch sharding and the
single flat-buffer gradient all-reduce.  The reference has no distributed code;
this is the multi-GPU row of SURVEY.md 8e.  Backend `nccl` on GPUs (NVLink 5 /
NVSwitch), `gloo` in the CPU tests."""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Strided shard of a dataset of n sequences (every sequence lands on exactly one rank)."""
    return np.arange(rank, n, world)


def shard_sequence_chunks(G: int, rank: int, world: int, chunk: int = 32) -> List[int]:
    """Generation shard of rank `rank`: the sequences of the predict chunks rank, rank+world, ... (chunk = Keras'
    predict batch of 32, which scopes the pitch_bins scramble of model.py:43-49, so a sharded run computes exactly
    what one GPU walking all chunks computes).  Every sequence lands on exactly one rank."""
    return [g for g in range(G) if (g // chunk) % world == rank]


def allreduce_flat(flat: torch.Tensor) -> torch.Tensor:
    """Sum the flat gradient buffer over ranks in place (one collective per step).
    The 1/world average is applied by the Nadam kernel (gscale)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def sum_over_ranks(t: torch.Tensor) -> torch.Tensor:
    """Sum a small tensor over ranks in place (epoch statistics)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class _DeviceBlock:
    """A device allocation that is not torch's, exposed through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr, False), version=2)


class PeerNadam:
    """Gradient exchange fused with the optimizer over peer memory (dj_nadam_allreduce_peer): the engine's flat
    parameter and gradient buffers move into a CUDA-IPC-shared allocation, every rank maps every other rank's
    allocation, and one kernel per step reduces this rank's slice of the gradient straight out of the peers'
    buffers, applies Nadam to it and stores the new weights into all of them.  Replaces
    `allreduce_flat(gflat); nadam_step(1/world)`; one node, one process per GPU, `torch.distributed` only for the
    handle exchange.  The Nadam moments of a slice exist on its owner only."""

    def __init__(self, eng):
        from . import _lib
        self.lib = _lib.load()
        self.check = _lib.check
        dbg = os.environ.get("DJ_PEER_DEBUG", "") != ""

        def stage(msg):
            if dbg:
                print(f"[PeerNadam rank {self.rank}] {msg}", file=sys.stderr, flush=True)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        n = eng.flat_size
        self.n = n
        self.flag_words = int(self.lib.dj_peer_flag_words())
        nbytes = (2 * n + self.flag_words) * 4
        # Setting up is a COLLECTIVE decision: a rank whose allocation or whose mapping of one peer fails must not
        # leave the others waiting in a collective (or running the peer kernel alone), so after each phase the
        # ranks exchange success flags and either all go on or all release what they hold and raise.
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        rc = self.lib.dj_peer_alloc(nbytes, C.byref(own), handle)
        err = None if rc == 0 else self.lib.dj_last_error().decode("utf-8", "replace")
        stage("allocated" if rc == 0 else f"allocation failed: {err}")
        gathered: List[Optional[tuple]] = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(gathered, (rc == 0, handle.raw, err))
        else:
            gathered[0] = (rc == 0, handle.raw, err)
        if not all(g[0] for g in gathered):
            if rc == 0:
                self.lib.dj_peer_free(own)
            bad = {r: g[2] for r, g in enumerate(gathered) if not g[0]}
            raise RuntimeError(f"dj_peer_alloc failed on rank(s) {bad}")
        handles = [g[1] for g in gathered]
        stage("handles exchanged")
        self.bases: List[int] = []
        opened, err = [], None
        for r in range(self.world):
            if r == self.rank:
                self.bases.append(own.value)
                continue
            q = C.c_void_p()
            rc = self.lib.dj_peer_open(handles[r], C.byref(q))
            if rc != 0:
                err = f"dj_peer_open(rank {r}): " + self.lib.dj_last_error().decode("utf-8", "replace")
                break
            opened.append(q.value)
            self.bases.append(q.value)
        oks: List[Optional[tuple]] = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(oks, (err is None, err))
        else:
            oks[0] = (err is None, err)
        if not all(o[0] for o in oks):
            for b in opened:
                self.lib.dj_peer_close(C.c_void_p(b))
            if self.world > 1:
                dist.barrier()            # nobody frees a block a peer still has mapped
            self.lib.dj_peer_free(own)
            raise RuntimeError(f"peer mapping failed: {[o[1] for o in oks if not o[0]]}")
        stage("peers opened")
        base = own.value
        self.flat = torch.as_tensor(_DeviceBlock(base, n, "<f4"), device=eng.dev)
        self.gflat = torch.as_tensor(_DeviceBlock(base + 4 * n, n, "<f4"), device=eng.dev)
        self.flags = torch.as_tensor(_DeviceBlock(base + 8 * n, self.flag_words, "<i4"), device=eng.dev)
        vp = C.c_void_p * self.world
        self._p = vp(*[b for b in self.bases])
        self._g = vp(*[b + 4 * n for b in self.bases])
        self._f = vp(*[b + 8 * n for b in self.bases])
        self.epoch = 0
        self.closed = False
        eng.rebind_flat(self.flat, self.gflat)
        eng.peer = self
        torch.cuda.synchronize()
        stage("weights moved")
        if self.world > 1:
            dist.barrier()       # every rank's buffers hold its weights before anybody's first step

    def next_epoch(self) -> int:
        self.epoch = self.epoch % 0xFFFFFFFF + 1
        return self.epoch

    def step_dev(self, eng, epoch_ptr: int, scalar_ptr: int, stream) -> None:
        """The same exchange with the epoch and the Nadam scalars read from device memory (Engine.StepParams): the
        form a captured CUDA graph of the training step replays.  The caller advances the epoch (next_epoch)."""
        eng._call("dj_nadam_allreduce_peer_dev", self._p, self._g, self._f, self.rank, self.world, eng.m.data_ptr(),
                  eng.v.data_ptr(), self.n, C.c_void_p(epoch_ptr), C.c_void_p(scalar_ptr), stream)

    def step(self, eng, scalars, stream) -> None:
        """One fused exchange + Nadam update; `scalars` = Engine._nadam_scalars()."""
        self.epoch = self.epoch % 0xFFFFFFFF + 1
        eng._call("dj_nadam_allreduce_peer", self._p, self._g, self._f, self.rank, self.world, eng.m.data_ptr(),
                  eng.v.data_ptr(), self.n, self.epoch, 1.0 / self.world, *scalars, stream)

    def status_word(self) -> torch.Tensor:
        """Device view of the sticky status word (0 = healthy, 1 = a bounded wait gave up); no synchronisation."""
        return self.flags[self.flag_words - 1]

    def raise_if_timed_out(self) -> None:
        """Host check (synchronises) of the status word a bounded wait sets when a peer never showed up.  From that
        launch on the kernel skips every update, so the weights are those of the last good step."""
        if int(self.status_word().item()) != 0:
            raise RuntimeError("dj_nadam_allreduce_peer: a rank did not reach the exchange within the time limit; "
                               "training stopped updating at that step")

    def close(self, eng=None) -> None:
        if self.closed:
            return
        torch.cuda.synchronize()
        if eng is not None and eng.peer is self:
            eng.rebind_flat(torch.empty_like(self.flat), torch.empty_like(self.gflat))
            eng.peer = None
        if self.world > 1:
            dist.barrier()
        for r, b in enumerate(self.bases):
            if r != self.rank:
                self.check(self.lib.dj_peer_close(C.c_void_p(b)), "dj_peer_close")
        if self.world > 1:
            dist.barrier()
        del self.flat, self.gflat, self.flags
        self.check(self.lib.dj_peer_free(C.c_void_p(self.bases[self.rank])), "dj_peer_free")
        self.closed = True


def make_step_exchange(eng, world: int):
    """What train_step should use at this world size: (allreduce callable or None, PeerNadam or None).
    On GPUs the default is the fused peer-memory kernel (measured on B200s: 30 us against 49 us for NCCL all-reduce
    + Nadam kernel at 2 ranks, 39 us against 65 us at 8); DJ_PEER_NADAM=0, or a node whose GPUs cannot map each
    other's memory, selects the NCCL all-reduce followed by dj_nadam_step."""
    if world <= 1:
        return None, None
    if os.environ.get("DJ_PEER_NADAM", "1") != "0" and torch.cuda.is_available():
        try:
            return None, PeerNadam(eng)
        except RuntimeError as e:      # raised on EVERY rank or on none (PeerNadam.__init__ decides collectively)
            print(f"[deepj] peer-memory exchange unavailable ({e}); using NCCL all-reduce", file=sys.stderr)
    return allreduce_flat, None
