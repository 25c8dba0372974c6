"""Hash-sharded repeat-graph construction across the GPUs of one node: one process per GPU, torch.distributed for
the plumbing (NCCL over NVLink / NVSwitch; gloo in the CPU tests).

The reference has no distributed mode; this is the multi-GPU form of the hot path that replaces
nrpcalc/base/hgraph.py:23-94:

  * parts are block-sharded: rank r extracts the k-mers of a contiguous range of parts (balanced by windows);
  * the k-mer space is hash-partitioned: owner(kmer) = owner_hash(kmer) % world, so all records of one k-mer meet
    on one rank -- ONE all-to-all of (k-mer, part) records;
  * a part's internal repeat is detected by whichever rank owns the repeated k-mer, but every rank must drop that
    part: ONE all-reduce (MAX) of the per-part flag bytes;
  * groups, their ranks, last single windows and edges are small and are gathered to every rank, so that each
    rank ends with the same RepeatStructure as the single-GPU path and runs the same host replay.

Every rank stages ALL parts on its device (record values are global part indices; first-window ranks can be
computed by the owner of a group without another exchange).

`ops` is the per-rank device back end: NativeOps drives libnrpb200.so; the CPU tests plug in a numpy model of the
same five calls to exercise the sharding / exchange / merge logic under gloo.
"""
import os
import weakref

import numpy as np
import torch
import torch.distributed as dist

# the record exchange is point-to-point traffic: let NCCL use many channels per peer (NVSwitch gives every pair full bandwidth)
os.environ.setdefault('NCCL_MIN_P2P_NCHANNELS', '32')
os.environ.setdefault('NCCL_MAX_P2P_NCHANNELS', '64')


def shard_bounds(lengths, k, world):
    """Contiguous part ranges, one per rank, balanced by the number of k-windows."""
    windows = np.maximum(np.asarray(lengths, dtype=np.int64) - k + 1, 0)
    total = int(windows.sum())
    csum = np.concatenate(([0], np.cumsum(windows)))
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world, side='left')))
    cuts.append(len(lengths))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(lengths)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


class _DeviceArray(object):
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


class NativeOps(object):
    """The five device calls of the sharded path on one GPU (include/nrp_b200.h, 'sharded path')."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device('cuda', ctx.device)
        self.stage_ms = {}

    def _account(self):
        for name, value in self.ctx.timings().items():
            self.stage_ms[name] = value           # cumulative inside the library for the current job

    def load(self, ascii_buf, offsets, ids, device_ptr=None):
        """Stage ALL parts on this rank's device."""
        self.ctx.load_parts(ascii_buf, offsets, ids, device_ptr=device_ptr)
        self.n_parts = len(offsets) - 1
        self.max_len = int(np.diff(offsets).max()) if self.n_parts else 0
        self.staged_bounds = None

    def load_shard(self, ascii_buf, offsets, ids, k, group=None, part_len=None, chunks=4):
        """COLLECTIVE.  Stage only this rank's shard (the part range build_sharded will give it): 1/world of the host-to-device
        traffic.  The device buffer, offsets, ids and record values stay global; the part buffers of all ranks are mapped
        through CUDA IPC so that the first window of a shared k-mer can be read from its home rank over NVLink.
        offsets=None with part_len: parts of one length (no offsets array is built or copied)."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if offsets is None:
            n = int(len(ids)) if ids is not None else int(len(ascii_buf) // part_len)
            cuts = [(n * r) // world for r in range(world + 1)]
            bounds = [(cuts[r], cuts[r + 1]) for r in range(world)]
            total, self.max_len = n * int(part_len), int(part_len)
        else:
            offsets = np.asarray(offsets, dtype=np.int64)
            n = offsets.size - 1
            lengths = np.diff(offsets)
            bounds = shard_bounds(lengths, k, world)
            total, self.max_len = int(offsets[-1]), (int(lengths.max()) if n else 0)
        needed = ((total + 4095) // 4096) * 4096 + 256
        if needed > self.ctx.parts_capacity():
            self._close_peers(group, force=True)           # the part buffer will be re-allocated: nobody may still map it
        lo, hi = bounds[rank]
        self.ctx.load_parts_shard(ascii_buf, offsets, part_len, ids, n, lo, hi, chunks=chunks)
        self.n_parts = n
        self.staged_bounds = bounds
        # exchange the IPC handles of the part buffers (64 bytes per rank) and open the peers'
        mine = torch.from_numpy(self.ctx.peer_export_parts()).to(self.device)
        everyone = torch.empty(64 * world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        self.ctx.peer_open_parts(everyone.cpu().numpy(), rank, [b[0] for b in bounds] + [n])
        self._parts_open = True

    def extract(self, lo, hi, k, internal_repeats, world):
        counts = self.ctx.shard_extract(lo, hi, k, internal_repeats, world)
        keys_ptr, vals_ptr, key_bytes = self.ctx.shard_send_buffers()
        n = int(counts.sum())
        self.key_dtype = torch.int64 if key_bytes == 8 else torch.int32
        if n == 0:
            return (torch.empty(0, dtype=self.key_dtype, device=self.device), torch.empty(0, dtype=torch.int32, device=self.device), counts)
        keys = torch.as_tensor(_DeviceArray(keys_ptr, n, '<i8' if key_bytes == 8 else '<i4'), device=self.device)
        vals = torch.as_tensor(_DeviceArray(vals_ptr, n, '<i4'), device=self.device)
        return keys, vals, counts

    # ---- fused exchange: records are stored straight into the owners' receive buffers over NVLink ------------------
    def peers_available(self):
        mode = os.environ.get('NRPCALC_B200_EXCHANGE', 'peer')
        if mode != 'peer' or torch.cuda.device_count() < 2:
            return False
        return True

    def _close_peers(self, group, force=False):
        """Before any rank re-exports (and possibly re-allocates) a buffer its peers map: every rank drops its mappings of the
        peers' buffers, then all meet at a barrier (freeing an exported allocation a peer still maps is undefined in CUDA)."""
        if (getattr(self, '_peer_key_direct', None) is None and getattr(self, '_peer_key', None) is None
                and not getattr(self, '_parts_open', False) and not force):
            return
        self.ctx.peer_close()
        self._peer_key_direct = self._peer_key = None
        self._parts_open = False
        dist.barrier(group=group)

    def exchange_direct(self, lo, hi, k, internal_repeats, windows_total, group, phases):
        """Direct fused exchange: no count pass.  Every (level-1 bucket, sender) pair owns a fixed region of the owner's
        receive buffer; scatter -> all-reduce(MAX) of the overflow flags (= barrier) -> the owner's level-2 split, filter and
        ordering run straight from the received regions.  Returns (sent, received) or None when a region overflowed on some
        rank (heavily duplicated input): the caller then uses the counted exchange."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if os.environ.get('NRPCALC_B200_DIRECT', '1') == '0':
            return None
        key = ('direct', int(windows_total), int(k), self.n_parts, world, int(getattr(self, 'max_len', 0)))
        if getattr(self, '_peer_key_direct', None) != key:
            self._close_peers(group)
            handle = self.ctx.peer_export_direct(windows_total, world, k)
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            self.ctx.peer_open_direct(handles, rank)
            self._peer_key_direct = key
            self._overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        phases.mark('peer_setup')
        overflow = self.ctx.shard_scatter_direct(lo, hi, k, internal_repeats, world)
        phases.mark('scatter')
        self._overflow.fill_(1 if overflow else 0)
        dist.all_reduce(self._overflow, op=dist.ReduceOp.MAX, group=group)      # also the barrier: every peer's stores have landed
        if int(self._overflow.item()) != 0:
            phases.mark('barrier')
            return None
        phases.mark('barrier')
        self.ctx.shard_group_direct(k, internal_repeats)
        return self.ctx.shard_moved()

    def exchange_fused(self, lo, hi, k, internal_repeats, windows_total, group, phases):
        """flags + counts -> all-gather of the counts -> peer-store scatter -> barrier -> group the received records.
        Returns (sent, received) or None when the receive capacity would overflow (caller falls back to NCCL)."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        capacity = int(windows_total / world * 1.15) + 65536
        key = (capacity, k <= 16, self.n_parts, world)
        if getattr(self, '_peer_key', None) != key:
            self._close_peers(group)
            handle = self.ctx.peer_export(capacity, k)
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            self.ctx.peer_open(handles, rank)
            self._peer_key = key
        phases.mark('peer_setup')
        counts = self.ctx.shard_count(lo, hi, k, internal_repeats, world)
        phases.mark('count_kernel')
        mine = torch.as_tensor(counts, device=self.device)
        matrix = torch.empty(world * world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(matrix, mine, group=group)      # also orders this build after every rank's previous one
        matrix = matrix.cpu().numpy().reshape(world, world)          # [sender, owner]
        phases.mark('counts')
        received = matrix.sum(axis=0)
        if int(received.max()) > capacity:
            return None
        offsets = matrix[:rank].sum(axis=0)
        self.ctx.shard_scatter_peers(k, offsets)
        phases.mark('scatter')
        dist.barrier(group=group)                                    # every peer's stores have landed
        phases.mark('barrier')
        self.ctx.shard_group_received(int(received[rank]), k, internal_repeats)
        return int(counts.sum()), int(received[rank])

    def group(self, keys, vals, k, internal_repeats):
        torch.cuda.current_stream(self.device).synchronize()
        self.ctx.shard_group(keys.data_ptr() if keys.numel() else 0, vals.data_ptr() if vals.numel() else 0, keys.numel(), k, internal_repeats)

    def exchange_repeat_flags(self, group):
        """Sparse form of the flag exchange: the parts this rank flagged for an internal repeat travel as a short list (one small
        all-gather instead of an all-reduce over n_parts bytes); finish() reports a truncated list."""
        world = dist.get_world_size(group)
        cap = int(os.environ.get('NRPCALC_B200_FLAG_CAP', '16384'))
        if getattr(self, '_flag_lists', None) is None or self._flag_lists[0].numel() != cap + 1 or self._flag_lists[1].numel() != world * (cap + 1):
            self._flag_lists = (torch.empty(cap + 1, dtype=torch.int32, device=self.device),
                                torch.empty(world * (cap + 1), dtype=torch.int32, device=self.device))
        mine, everyone = self._flag_lists
        self.ctx.flags_repeat_list(mine.data_ptr(), cap)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        torch.cuda.current_stream(self.device).synchronize()
        self.ctx.flags_repeat_apply(everyone.data_ptr(), world, cap)

    def flags(self):
        out = torch.empty(max(self.n_parts, 1), dtype=torch.uint8, device=self.device)
        self.ctx.flags_export(out.data_ptr())
        return out[:self.n_parts]

    def set_flags(self, flags):
        torch.cuda.current_stream(self.device).synchronize()
        if self.n_parts:
            self.ctx.flags_import(flags.data_ptr())

    def _view(self, what, dtype, typestr):
        ptr, n = self.ctx.result_device(what)
        if n == 0:
            return torch.empty(0, dtype=dtype, device=self.device)
        return torch.as_tensor(_DeviceArray(ptr, n, typestr), device=self.device)

    def finish(self, k, internal_repeats, want_edges):
        """Groups of this rank stay on the device; returns zero-copy views of (group keys, edge u, edge v)."""
        if not self.ctx.shard_finish(k, internal_repeats, want_edges):
            return None                                   # truncated flag list (see exchange_repeat_flags)
        keys = self._view(0, torch.int64, '<i8')
        if want_edges:
            return keys, self._view(1, torch.int32, '<i4'), self._view(2, torch.int32, '<i4')
        return keys, None, None

    def last_single(self, all_keys, k):
        torch.cuda.current_stream(self.device).synchronize()
        self.ctx.shard_last_single(None, k, device_ptr=all_keys.data_ptr() if all_keys.numel() else 0, n=all_keys.numel())

    def merge_edges(self, codes):
        """codes: int64 tensor of (u << 32) | v gathered from all ranks"""
        torch.cuda.current_stream(self.device).synchronize()
        self.ctx.edges_merge(codes.data_ptr() if codes.numel() else 0, 0, codes.numel())

    def to_host(self, tensor, pool='default'):
        """Device tensor -> numpy VIEW of a cached page-locked staging buffer (a pageable .cpu() runs at a fraction of PCIe speed;
        a second copy out of the staging buffer would cost as much again).  One buffer per `pool` name; the view is valid until the
        next call with the same name."""
        nbytes = tensor.numel() * tensor.element_size()
        if nbytes == 0:
            return np.zeros(0, dtype=torch.empty(0, dtype=tensor.dtype).numpy().dtype)
        view = self._pinned(pool, nbytes)[:nbytes].view(tensor.dtype)
        view.copy_(tensor.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return view.numpy()

    def _pinned(self, pool, nbytes):
        pools = self.__dict__.setdefault('_pinned_pools', {})
        buf = pools.get(pool)
        if buf is None or buf.numel() < nbytes:
            buf = pools[pool] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
        return buf

    def gather_results(self, group, lo, hi, n, want_edges):
        """Every rank's groups and last single windows on every host, through NCCL all-gathers of the arrays themselves (device
        views of the library's result buffers): one small size exchange, one padded all-gather of the group arrays, one of the
        last-single slices whose rows are copied straight to their place in a page-locked result array.  The returned arrays are
        views of page-locked buffers owned by this object (valid until the next call)."""
        world = dist.get_world_size(group)
        counts = self.ctx.counts()
        indptr = self._view(5, torch.int64, '<i8')
        members = self._view(6, torch.int32, '<i4')
        first_window = self._view(7, torch.int32, '<i4')
        keys = self._view(0, torch.int64, '<i8')
        last_single = self._view(4, torch.int32, '<i4')[lo:hi] if n else torch.empty(0, dtype=torch.int32, device=self.device)
        ng, nm = int(keys.numel()), int(members.numel())
        sizes = torch.tensor([ng, nm, hi - lo, lo, counts['candidates'], counts['oversized']], dtype=torch.int64, device=self.device)
        all_sizes = torch.empty(6 * world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(all_sizes, sizes, group=group)
        all_sizes = all_sizes.cpu().view(world, 6).tolist()
        # group arrays (small): one int32 payload per rank = group sizes | members | first windows | keys (two words each)
        group_sizes = (indptr[1:] - indptr[:-1]).to(torch.int32) if ng else torch.empty(0, dtype=torch.int32, device=self.device)
        payload = torch.cat([group_sizes, members, first_window, keys.view(torch.int32)])
        longest = max(4 * s[0] + s[1] for s in all_sizes)
        padded = torch.empty(max(longest, 1), dtype=torch.int32, device=self.device)
        padded[:payload.numel()] = payload
        everything = torch.empty(world * padded.numel(), dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(everything, padded, group=group)
        host = self.to_host(everything, 'groups').reshape(world, -1)
        parts = {'sizes': [], 'members': [], 'first_window': [], 'keys': []}
        for r, (g, m, _, _, _, _) in enumerate(all_sizes):
            row, at = host[r], 0
            for name, count in (('sizes', g), ('members', m), ('first_window', g), ('keys', 2 * g)):
                parts[name].append(row[at:at + count])
                at += count
        indptr_all = np.zeros(sum(s[0] for s in all_sizes) + 1, dtype=np.int64)
        np.cumsum(np.concatenate(parts['sizes']).astype(np.int64), out=indptr_all[1:])
        # last single windows (n entries in all): padded rows on the device, each row copied to its place on the host
        last_all = np.zeros(0, dtype=np.int32)
        if n:
            widest = max(s[2] for s in all_sizes)
            row = torch.empty(max(widest, 1), dtype=torch.int32, device=self.device)
            row[:last_single.numel()] = last_single
            rows = torch.empty(world * row.numel(), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(rows, row, group=group)
            target = self._pinned('last_single', 4 * n)[:4 * n].view(torch.int32)
            for r, (_, _, span, first, _, _) in enumerate(all_sizes):
                if span:
                    target[first:first + span].copy_(rows[r * row.numel():r * row.numel() + span], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            last_all = target.numpy()
        edges = None
        if want_edges:
            eu, ev = self._view(1, torch.int32, '<i4'), self._view(2, torch.int32, '<i4')
            edges = (self.to_host(eu, 'edge_u').view(np.uint32), self.to_host(ev, 'edge_v').view(np.uint32))
        return {'indptr': indptr_all, 'members': np.concatenate(parts['members']).view(np.uint32),
                'first_window': np.concatenate(parts['first_window']).view(np.uint32),
                'keys': np.ascontiguousarray(np.concatenate(parts['keys'])).view(np.uint64),
                'last_single': last_all, 'edges': edges, 'candidates': [s[4] for s in all_sizes], 'oversized': [s[5] for s in all_sizes]}

    def fetch_host(self, want_edges):
        """Host copies of this rank's groups, the flags, this shard's last single windows and the merged edges."""
        counts = self.ctx.counts()
        res = self.ctx.fetch(want_keys=True)
        return {'indptr': res.indptr, 'members': res.members, 'first_window': res.first_window, 'keys': res.keys,
                'last_single': res.last_single, 'flags': res.flags,
                'edges': (res.edges[0].copy(), res.edges[1].copy()) if want_edges and res.edges is not None else None,
                'candidates': counts['candidates'], 'oversized': counts['oversized']}


def _all_gather_var(tensor, group, extra=0):
    """all-gather of 1-D tensors of different lengths (pads to the longest, NCCL needs equal sizes).
    Returns (concatenation in rank order, [(length, extra) per rank]); `extra` is a small integer shipped along."""
    world = dist.get_world_size(group)
    n = torch.tensor([tensor.numel(), int(extra)], dtype=torch.int64, device=tensor.device)
    sizes = torch.empty(2 * world, dtype=torch.int64, device=tensor.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = sizes.cpu().view(world, 2).tolist()
    longest = max(s[0] for s in sizes) if sizes else 0
    if longest == 0:
        return tensor.new_empty(0), [(0, 0)] * world
    padded = tensor.new_empty(longest)
    padded[:tensor.numel()] = tensor
    out = tensor.new_empty(world * longest)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * longest:r * longest + sizes[r][0]] for r in range(world)]), [(int(a), int(b)) for a, b in sizes]


def _all_to_all_records(keys, vals, send_counts, group, phases=None):
    """One all-to-all of the owner-partitioned records.  Returns (recv_keys, recv_vals, recv_counts)."""
    world = dist.get_world_size(group)
    device = keys.device
    send = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, group=group)
    recv_counts = recv.cpu().numpy()
    if phases is not None:
        phases.mark('a2a_counts')
    n_recv = int(recv_counts.sum())
    recv_keys = torch.empty(n_recv, dtype=keys.dtype, device=device)
    recv_vals = torch.empty(n_recv, dtype=vals.dtype, device=device)
    in_splits = [int(x) for x in send_counts]
    out_splits = [int(x) for x in recv_counts]
    if phases is not None:
        phases.mark('a2a_alloc')
    dist.all_to_all_single(recv_keys, keys, out_splits, in_splits, group=group)
    if phases is not None:
        phases.mark('a2a_keys')
    dist.all_to_all_single(recv_vals, vals, out_splits, in_splits, group=group)
    if phases is not None:
        phases.mark('a2a_vals')
    return recv_keys, recv_vals, recv_counts


class _Phases(object):
    """Wall-clock per phase (device work is synchronised at every phase boundary by the calls themselves)."""

    def __init__(self, sink, device):
        import time
        self.sink, self.device, self.clock = sink, device, time.perf_counter
        self.t = self.clock()

    def mark(self, name):
        if self.sink is None:
            return
        if self.device.type == 'cuda':
            torch.cuda.synchronize(self.device)
        now = self.clock()
        self.sink[name] = self.sink.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


def build_sharded(ops, offsets, k, internal_repeats, want_edges=False, group=None, phase_ms=None, gather_host=True):
    """Run the sharded build over the parts already staged in `ops`.

    Device phase (what the graph-build metric times): extract -> all-to-all -> group -> flag OR-reduce -> finish
    (this rank's groups + first windows) -> all-gather of the shared k-mers -> last single windows of the shard ->
    all-gather + dedupe of the edges.  Afterwards flags and edges are complete on every device, groups and last
    single windows are distributed.

    Host phase (gather_host=True): every rank collects all groups and last single windows and returns the same
    dict: flags[n], indptr, members (global part indices), first_window, keys, last_single[n], edges (u, v) or
    None, plus statistics under 'stats'.  With gather_host=False only 'stats' is returned."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    plan = getattr(ops, '_shard_plan', None)
    stamp = (n, k, world, int(offsets[-1]) if n else 0, int(offsets[n // 2]) if n else 0)
    if plan is None or plan[0] != stamp or plan[4]() is not offsets:
        windows = np.maximum(np.diff(offsets) - k + 1, 0)
        # host-side plan of a staged batch, reused by repeated builds over the SAME (still alive) offsets array
        plan = (stamp, shard_bounds(np.diff(offsets), k, world), windows, int(windows.sum()), weakref.ref(offsets))
        ops._shard_plan = plan
    bounds, windows, windows_total = plan[1], plan[2], plan[3]
    staged = getattr(ops, 'staged_bounds', None)
    if staged is not None:                      # shard-only staging: the shards ARE what each rank holds
        bounds = staged
    lo, hi = bounds[rank]

    phases = _Phases(phase_ms, ops.device)
    moved = None
    if hasattr(ops, 'exchange_direct') and ops.peers_available():
        moved = ops.exchange_direct(lo, hi, k, internal_repeats, windows_total, group, phases)
    if moved is None and hasattr(ops, 'exchange_fused') and ops.peers_available():
        moved = ops.exchange_fused(lo, hi, k, internal_repeats, windows_total, group, phases)
    if moved is None:
        keys, vals, send_counts = ops.extract(lo, hi, k, internal_repeats, world)
        phases.mark('extract')
        recv_keys, recv_vals, recv_counts = _all_to_all_records(keys, vals, send_counts, group, phases if phase_ms is not None else None)
        phases.mark('all_to_all')
        ops.group(recv_keys, recv_vals, k, internal_repeats)
        moved = (int(np.sum(send_counts)), int(np.sum(recv_counts)))
    phases.mark('group')
    flags_all = None
    if hasattr(ops, 'exchange_repeat_flags') and os.environ.get('NRPCALC_B200_DENSE_FLAGS', '0') != '1':
        ops.exchange_repeat_flags(group)                # only the repeat bit has to reach the other ranks during the device phase
    else:
        flags_all = _or_reduce_flags(ops, ops.flags(), group)
        ops.set_flags(flags_all)
    phases.mark('flag_reduce')
    finished = ops.finish(k, internal_repeats, want_edges)
    if finished is None:                                # a rank's list of flagged parts was truncated: the dense all-reduce
        flags_all = _or_reduce_flags(ops, ops.flags(), group)
        ops.set_flags(flags_all)
        finished = ops.finish(k, internal_repeats, want_edges)
    group_keys, edge_u, edge_v = finished
    phases.mark('finish')
    # one variable-length all-gather carries both small results: the shared k-mers and the edge codes
    n_keys = group_keys.numel()
    if want_edges:
        codes = (edge_u.to(torch.int64) << 32) | (edge_v.to(torch.int64) & 0xffffffff)
        packed = torch.cat([group_keys, codes])
    else:
        packed = group_keys
    gathered_small, sizes_small = _all_gather_var(packed, group, extra=n_keys)
    all_keys, all_codes = [], []
    pos = 0
    for total, nk in sizes_small:
        all_keys.append(gathered_small[pos:pos + nk])
        all_codes.append(gathered_small[pos + nk:pos + total])
        pos += total
    ops.last_single(torch.cat(all_keys) if all_keys else group_keys, k)
    phases.mark('last_single')
    if want_edges:
        ops.merge_edges(torch.cat(all_codes))                       # a pair can come from several owners
    phases.mark('edges')

    stats = {'bounds': bounds, 'sent_here': moved[0], 'received_here': moved[1]}
    if not gather_host:
        return {'stats': stats}

    totals = torch.tensor([stats['sent_here'], stats['received_here']], dtype=torch.int64, device=ops.device)
    per_rank = torch.empty(2 * world, dtype=torch.int64, device=ops.device)
    dist.all_gather_into_tensor(per_rank, totals, group=group)
    per_rank = per_rank.cpu().numpy().reshape(world, 2)
    stats['sent'] = per_rank[:, 0].tolist()
    stats['received'] = per_rank[:, 1].tolist()
    if flags_all is None:                                # sparse exchange: the other ranks' palindrome / background bits are still missing
        flags_all = _or_reduce_flags(ops, ops.flags(), group)
        ops.set_flags(flags_all)
    flags_host = (ops.to_host(flags_all, 'flags') if hasattr(ops, 'to_host') else flags_all.cpu().numpy()) if n else np.zeros(0, dtype=np.uint8)
    # records of parts flagged for an internal repeat were exchanged before the flag was known (single path: same); the few
    # flagged parts are picked out on the device
    repeated = torch.nonzero(flags_all & 1).reshape(-1).cpu().numpy() if n else np.zeros(0, dtype=np.int64)
    stats['records'] = int(per_rank[:, 1].sum()) - int(windows[repeated].sum())
    phases.mark('host_flags')

    if hasattr(ops, 'gather_results'):
        result = ops.gather_results(group, lo, hi, n, want_edges)
        stats['candidates'], stats['oversized'] = result.pop('candidates'), result.pop('oversized')
        result.update({'flags': flags_host, 'stats': stats})
    else:
        local = ops.fetch_host(want_edges)
        gathered = [None] * world
        dist.all_gather_object(gathered, {'indptr': local['indptr'], 'members': local['members'], 'first_window': local['first_window'],
                                          'keys': local['keys'], 'last_single': local['last_single'][lo:hi],
                                          'candidates': local['candidates'], 'oversized': local['oversized']}, group=group)
        indptr = [np.zeros(1, dtype=np.int64)]
        base = 0
        for g in gathered:
            indptr.append(g['indptr'][1:] + base)
            base += int(g['indptr'][-1])
        stats['candidates'] = [g['candidates'] for g in gathered]
        stats['oversized'] = [g['oversized'] for g in gathered]
        result = {
            'flags': flags_host,
            'indptr': np.concatenate(indptr),
            'members': np.concatenate([g['members'] for g in gathered]),
            'first_window': np.concatenate([g['first_window'] for g in gathered]),
            'keys': np.concatenate([g['keys'] for g in gathered]),
            'last_single': np.concatenate([g['last_single'] for g in gathered]) if n else np.zeros(0, dtype=np.int32),
            'edges': local['edges'],
            'stats': stats,
        }
    phases.mark('host_gather')
    return result


def _or_reduce_flags(ops, flags_dev, group):
    """Combine the per-part flag bytes of all ranks with ONE all-reduce(MAX).
    MAX equals OR here: the palindrome / background bits of a part are set only by the rank whose shard holds the
    part, and such a part sends no records, so no other rank can add the repeat bit (value 1) to it; for every
    other part the only possible values are 0 and 1."""
    if flags_dev.numel():
        dist.all_reduce(flags_dev, op=dist.ReduceOp.MAX, group=group)
    return flags_dev
