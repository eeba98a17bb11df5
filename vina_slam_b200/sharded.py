"""Voxel map sharded by hash range across GPUs (BASELINE.json configs[4]: "a voxel map sharded by
hash range across 1/2/4/8 B200 over NVLink").

One process per GPU, one `capi.Ctx` per process holding the shard of the map whose root voxels hash into the
rank's range (`vina_shard_owner`). Per scan every rank
  1. routes its slice of the scan's down-sampled pointVar set (`vina_shard_route`: pvec_update + key + owner,
     stable partition by owner, on the device),
  2. exchanges the records with one all-to-all (`torch.distributed.all_to_all_single` over NCCL / NVLink - the
     only collective on the data path; PyTorch is plumbing here: buffers, streams, the process group),
  3. inserts what it received (`vina_shard_insert_begin/finish` around a 2-int all-reduce that makes the
     reference's "fewer roots than threads" early-outs global), then recut / margi locally.
The union of the shards is bit-identical to the map one GPU builds from the same scans (tests/test_gpu_parity.py,
tests/test_sharded_cpu.py for the exchange logic under gloo).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import math

import numpy as np

from . import capi

REC = capi.SHARD_RECORD_DOUBLES


def slice_of(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Ascending, contiguous share of n points for `rank` (first, count): the concatenation over ranks is the scan
    order, which is what makes the sharded sums bit-identical to the single-GPU ones."""
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def exchange_records(send, counts: Sequence[int], group=None):
    """All-to-all of routed records. send: (n, REC) float64 tensor grouped by destination rank (counts[r] rows for
    rank r, in that order). Returns the (m, REC) tensor of rows received, ordered by source rank. Works on CUDA
    tensors (nccl) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    counts = [int(c) for c in counts]
    world = len(counts)
    if world == 1 or not dist.is_initialized():
        return send[: counts[0]]
    c_out = torch.tensor(counts, dtype=torch.int64, device=send.device)
    c_in = torch.empty_like(c_out)
    dist.all_to_all_single(c_in, c_out, group=group)
    rc = [int(v) for v in c_in.cpu().tolist()]
    recv = torch.empty((sum(rc), REC), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send[: sum(counts)].contiguous(), output_split_sizes=rc, input_split_sizes=counts,
                           group=group)
    return recv


def local_exchange(sends: List, counts: List[Sequence[int]]):
    """The same permutation without a process group: all ranks' send buffers live in this process (several
    contexts on one GPU, or numpy arrays on the CPU). Returns the per-rank receive buffers."""
    world = len(sends)
    starts = [np.concatenate([[0], np.cumsum(np.asarray(c, dtype=np.int64))]) for c in counts]
    out = []
    for dst in range(world):
        parts = [sends[src][int(starts[src][dst]): int(starts[src][dst + 1])] for src in range(world)]
        if isinstance(parts[0], np.ndarray):
            out.append(np.concatenate(parts, axis=0))
        else:
            import torch

            out.append(torch.cat(parts, dim=0).contiguous())
    return out


class MapShard:
    """This rank's shard and the sliding-window bookkeeping of `map_update` (host/vina_pipeline.cpp;
    local_mapping.cpp:425-451, 489-546 with if_BA == 0)."""

    def __init__(self, ctx: capi.Ctx, rank: int, world: int, device=None, group=None, own_stream: bool = False):
        import torch

        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # one stream for the kernels of the ctx, torch's copies and the collectives (NCCL orders itself against it);
        # own_stream keeps the context's private stream instead - needed when several ranks live in one process
        # and exchange through peer memory (a rank waiting on the device must not block the others' kernels)
        if not own_stream:
            ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.win_count = 0
        self.win_base = 0
        self.x_buf: List[Tuple[np.ndarray, np.ndarray]] = []
        self._send = None
        # distance travelled / pruning (local_mapping.cpp:262-263, 272, 509-519): every rank sees every pose, so the
        # bookkeeping is replicated and each rank prunes its own shard - no collective
        self.jour, self.last_pos, self.release_flag = 0.0, (0.0, 0.0, 0.0), False

    # -- step 1
    def route(self, first: int, count: int, index_base: int, R_col, p, cov_rot_col, cov_tsl_col):
        import torch

        if self._send is None or self._send.shape[0] < max(count, 1):
            self._send = torch.empty((max(count, 1), REC), dtype=torch.float64, device=self.device)
        counts = self.ctx.shard_route(self.world, first, count, index_base, R_col, p, cov_rot_col, cov_tsl_col,
                                      self._send.data_ptr())
        return self._send[:count], counts

    # -- step 3a / 3b
    def insert_begin(self, recv):
        self._recv = recv  # keep the buffer alive until the kernels have consumed it
        return self.ctx.shard_insert_begin(recv.data_ptr() if recv.shape[0] else 0, int(recv.shape[0]), self.win_count - 1)

    def insert_finish(self, global_roots: int, global_slide: int):
        self.ctx.shard_insert_finish(self.win_count - 1, int(global_roots), int(global_slide))

    def push_pose(self, R_col, p):
        self.win_count += 1
        self.x_buf.append((np.array(R_col, dtype=np.float64), np.array(p, dtype=np.float64)))

    def recut_margi(self):
        xb = np.zeros(len(self.x_buf), dtype=capi.POSE_DTYPE)
        for i, (R, p) in enumerate(self.x_buf):
            xb[i]["R"], xb[i]["p"] = R, p
        self.ctx.map_recut(self.win_count, xb)
        if self.win_count >= self.ctx.cfg.win_size:
            self.ctx.map_set_journey(self.jour)  # the stamp of multi_margi (local_mapping.cpp:36, 507)
            self.ctx.map_margi(self.win_count, xb)
            if (self.win_base + self.win_count) % 10 == 0:
                p = [float(v) for v in self.x_buf[-1][1]]
                d = [p[i] - self.last_pos[i] for i in range(3)]
                spat = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
                if spat > 0.5:
                    self.jour += spat
                    self.last_pos = tuple(p)
                    self.release_flag = True
            self.x_buf.pop(0)
            self.win_base += 1
            self.win_count -= 1

    def idle(self, horizon: int = 700):
        """The idle path's map pruning on this rank's shard (local_mapping.cpp:317-341); (roots erased, nodes freed)."""
        if not self.release_flag:
            return 0, 0
        self.release_flag = False
        return self.ctx.map_prune(self.jour, horizon)

    # -- fused route + exchange over peer memory (NVLink): no staging, no collective call for the records
    def p2p_setup(self, inbox_records: int):
        """Allocate this rank's inbox and connect to the peers' (CUDA IPC handles all-gathered over the group)."""
        import torch.distributed as dist

        mine = self.ctx.shard_p2p_create(self.rank, self.world, inbox_records)
        if self.world > 1:
            blobs = [None] * self.world
            dist.all_gather_object(blobs, mine, group=self.group)
            self.ctx.shard_p2p_connect(b"".join(blobs))
        self.p2p = True

    def update_p2p(self, first: int, count: int, index_base: int, R_col, p, cov_rot_col, cov_tsl_col):
        import torch
        import torch.distributed as dist

        self.push_pose(R_col, p)
        self.ctx.shard_route_p2p(first, count, index_base, R_col, p, cov_rot_col, cov_tsl_col)
        n, roots, slide = self.ctx.shard_insert_begin_p2p(self.win_count - 1)
        tot = torch.tensor([roots, slide], dtype=torch.int64, device=self.device)
        if self.world > 1 and dist.is_initialized():
            dist.all_reduce(tot, group=self.group)  # also the barrier that frees the inboxes for the next scan
        g = tot.cpu().tolist()
        self.insert_finish(g[0], g[1])
        self.recut_margi()
        return n

    # -- the whole per-scan map update, collectives included
    def update(self, first: int, count: int, index_base: int, R_col, p, cov_rot_col, cov_tsl_col):
        import torch
        import torch.distributed as dist

        self.push_pose(R_col, p)
        send, counts = self.route(first, count, index_base, R_col, p, cov_rot_col, cov_tsl_col)
        recv = exchange_records(send, counts, self.group)
        roots, slide = self.insert_begin(recv)
        tot = torch.tensor([roots, slide], dtype=torch.int64, device=self.device)
        if self.world > 1 and dist.is_initialized():
            dist.all_reduce(tot, group=self.group)
        g = tot.cpu().tolist()
        self.insert_finish(g[0], g[1])
        self.recut_margi()
        return int(recv.shape[0])


def map_digest(nodes: np.ndarray) -> int:
    """Order-independent digest of exported octree nodes: sum of a CRC per node over every field. The digests of
    the shards add up to the digest of the single-GPU map exactly when the union equals it byte for byte."""
    import zlib

    if nodes.shape[0] == 0:
        return 0
    cols = [np.ascontiguousarray(nodes[f]).reshape(nodes.shape[0], -1).view(np.uint8) for f in nodes.dtype.names]
    raw = np.ascontiguousarray(np.concatenate(cols, axis=1))
    return int(sum(zlib.crc32(row.tobytes()) for row in raw))


QREC = capi.SHARD_QUERY_DOUBLES


def exchange_rows(send, counts: Sequence[int], width: int, group=None):
    """exchange_records for rows of `width` doubles (association queries are 10 wide, map records 13)."""
    import torch
    import torch.distributed as dist

    counts = [int(c) for c in counts]
    if len(counts) == 1 or not dist.is_initialized():
        return send[: counts[0]]
    c_out = torch.tensor(counts, dtype=torch.int64, device=send.device)
    c_in = torch.empty_like(c_out)
    dist.all_to_all_single(c_in, c_out, group=group)
    rc = [int(v) for v in c_in.cpu().tolist()]
    recv = torch.empty((sum(rc), width), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send[: sum(counts)].contiguous(), output_split_sizes=rc, input_split_sizes=counts,
                           group=group)
    return recv


class ShardedIekf:
    """LioStateEstimation (odometry.cpp:64-255) against a map sharded by hash range: per iteration every rank
    routes its slice of the scan to the owners of the voxels the points fall into (all-to-all of 80-byte
    pointVar records), the owners evaluate gate / residual / Jacobian against their shard, the 34 sums are
    all-reduced and every rank applies the same update (the reference's 15x15 route on the host)."""

    def __init__(self, shard: MapShard):
        import torch

        self.sh = shard
        self._send = None
        self._sums = torch.zeros(34, dtype=torch.float64, device=shard.device)

    def route(self, first: int, count: int, R_col, p):
        import torch

        if self._send is None or self._send.shape[0] < max(count, 1):
            self._send = torch.empty((max(count, 1), QREC), dtype=torch.float64, device=self.sh.device)
        counts = self.sh.ctx.shard_query_route(self.sh.world, first, count, 0, R_col, p, self._send.data_ptr())
        return self._send[:count], counts

    def accumulate(self, recv, R_col, p, rot_var_col, tsl_var_col):
        self._recv = recv
        self.sh.ctx.shard_query_accumulate(recv.data_ptr() if recv.shape[0] else 0, int(recv.shape[0]), R_col, p,
                                           rot_var_col, tsl_var_col, self._sums.data_ptr())
        return self._sums

    def run(self, first: int, count: int, max_iter: int = 4):
        """The whole loop with the collectives; the ctx holds x_curr (same on every rank) and the full scan's
        pointVar set (vina_var_init(ctx, 0)). Returns the number of iterations."""
        import torch
        import torch.distributed as dist

        ctx = self.sh.ctx
        s0 = capi.state_arrays(ctx.get_state())
        rv = np.ascontiguousarray(s0["cov"][0:3, 0:3].T.reshape(-1))
        tv = np.ascontiguousarray(s0["cov"][3:6, 3:6].T.reshape(-1))
        ctx.odom_iekf_host_begin(max_iter)
        for it in range(max_iter):
            s = capi.state_arrays(ctx.get_state())
            Rc = np.ascontiguousarray(s["R"].T.reshape(-1))
            send, counts = self.route(first, count, Rc, s["p"])
            recv = exchange_rows(send, counts, QREC, self.sh.group)
            sums = self.accumulate(recv, Rc, s["p"], rv, tv)
            if self.sh.world > 1 and dist.is_initialized():
                dist.all_reduce(sums, group=self.sh.group)
            if ctx.odom_iekf_host_update(sums.cpu().numpy()):
                return it + 1
        return max_iter

    def run_p2p(self, first: int, count: int, max_iter: int = 4):
        """The same loop with the exchange and the update fused into the kernels (vina_odom_iekf_sharded_p2p):
        queries stored into the owners' inboxes over peer memory, the 34 sums to every peer's control block, the
        update on every rank's device iterate - no collective call and no host synchronisation inside the loop.
        Needs MapShard.p2p_setup. Returns the number of iterations."""
        it, _ = self.sh.ctx.odom_iekf_sharded_p2p(first, count, max_iter, capi.SHARD_IEKF_ALL)
        return it
