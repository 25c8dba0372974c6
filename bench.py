#!/usr/bin/env python
"""Finder graph-build benchmark (BASELINE.json metric: k-mers/s, parts/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2]

A step = one build of the repeat structure of one batch of synthetic random-DNA parts (BASELINE config 2 by
default: 1M x 100 bp, Lmax=16, internal_repeats=False).

  value     k-mers/s with the parts resident in HBM, from the library's CUDA events around the pipeline
            (device-resident inputs -> device-resident flags/groups/ranks/edges), summed over the K steps
  e2e       the same through the C ABI with HOST buffers: nrp_load_parts (H2D) + nrp_build + all result getters
            (D2H), wall clock around the calls
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration against the measured HBM copy peak
  cpu_baseline  oracle/nrp_oracle.c (hash-table restatement of the reference's algorithm) on the host cores

  parity    the full result of the last build (flags, groups with ranks, last unshared windows, edges) against
            oracle/nrp_oracle.c on the SAME full-size batch (default for c1/c2; --verify forces it for c3/c4/dup,
            --no-verify skips it); a mismatch makes the run exit non-zero

--impl reference times the UNMODIFIED reference (hgraph.get_homology_graph of the offline install under baseline/_ref,
or /root/reference in the development container) on a bounded sample of the same workload, one core like the reference;
without a reference tree it times the literal pure-Python port (oracle/ref_port.py) and says so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CONFIGS = {
    # name: (n_parts, length or (lo, hi), Lmax, internal_repeats, background genome bp, seed)
    'c1': dict(n=10_000, length=100, Lmax=12, internal_repeats=False, genome=0, seed=1,
               label='Finder mode: 10k synthetic 100-bp parts, Lmax=12, no background'),
    'c2': dict(n=1_000_000, length=100, Lmax=16, internal_repeats=False, genome=0, seed=2,
               label='Finder mode: 1M synthetic 100-bp parts, Lmax=16, internal_repeats=False, on 1 B200'),
    'c3': dict(n=1_000_000, length=200, Lmax=15, internal_repeats=False, genome=5_000_000, seed=3,
               label='Finder mode: 1M 200-bp parts with a 5-Mbp synthetic background genome k-mer set, Lmax=15'),
    'c4': dict(n=10_000_000, length=100, Lmax=20, internal_repeats=False, genome=0, seed=4,
               label='Finder mode: 10M synthetic 100-bp parts, Lmax=20, hash-sharded across 2/4/8 B200'),
    'c5': dict(n=50_000_000, length=(50, 300), Lmax=24, internal_repeats=False, genome=0, seed=5,
               label='Finder mode: 50M variable-length (50-300 bp) parts, Lmax=24, 8xB200 all-to-all scaling sweep'),
    # not a BASELINE shape: planted shared motifs, so that the fixed regions of the direct placement overflow and the exact
    # (histogram) path with the global sort runs -- the workload real part libraries resemble more than i.i.d. DNA
    'dup': dict(n=1_000_000, length=100, Lmax=16, internal_repeats=False, genome=0, seed=6, dup_motifs=4, dup_fraction=0.008,
                label='Finder mode: 1M 100-bp parts, Lmax=16, 0.8 % of the parts carry one of 4 shared 40-bp motifs (~2000 copies of each shared k-mer: duplicate-heavy)'),
}
# oracle passes over slices of the key space for the full-size verification (bounds the host memory of the C hash table)
VERIFY_SLICES = {'c1': 1, 'c2': 4, 'c3': 8, 'c4': 32, 'c5': 256, 'dup': 4}


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks and throttle reasons during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None
        self.first = 0          # rows before this index were taken before the timed region (warm-up)
        self.extra_steps = 0

    def mark(self):
        """The timed region starts now: only later samples are reported."""
        self.first = len(self.rows)

    def need_more(self, want=2):
        return self.proc is not None and len(self.rows) - self.first < want

    def top_up(self, step, want=2, limit_s=4.0):
        """A timed region of a few milliseconds can end before nvidia-smi has printed a row: keep the device under the SAME
        load with extra UNTIMED steps until `want` rows exist (reported as extra_untimed_steps)."""
        t0 = time.perf_counter()
        while self.need_more(want) and time.perf_counter() - t0 < limit_s:
            step()
            self.extra_steps += 1

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for row in self.rows[self.first:]:
            try:
                sm.append(float(row[0])); mx.append(float(row[1]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), row[3:7]):
                if cell.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'extra_untimed_steps': self.extra_steps}


def make_workload(cfg, n_parts=None):
    import cases
    n = cfg['n'] if n_parts is None else n_parts
    if isinstance(cfg['length'], tuple):
        return cases.random_varlen_buffer(n, cfg['length'][0], cfg['length'][1], cfg['seed'])
    buf, off = cases.random_dna_buffer(n, cfg['length'], cfg['seed'])
    if cfg.get('dup_motifs'):
        # plant one of a few shared motifs at a random position of a fraction of the parts (every shared window collides)
        rng = np.random.default_rng(cfg['seed'] + 1000)
        motif_len = 40
        motifs = cases.random_dna_buffer(cfg['dup_motifs'], motif_len, cfg['seed'] + 2000)[0].reshape(cfg['dup_motifs'], motif_len)
        buf = buf.copy()
        chosen = np.nonzero(rng.random(n) < cfg['dup_fraction'])[0]
        which = rng.integers(0, cfg['dup_motifs'], size=chosen.size)
        at = rng.integers(0, cfg['length'] - motif_len + 1, size=chosen.size)
        view = buf.reshape(n, cfg['length'])
        cols = at[:, None] + np.arange(motif_len)[None, :]
        view[chosen[:, None], cols] = motifs[which]
    return buf, off


def background_keys(cfg, ctx=None, timing=None):
    """Canonical K-mer set of the synthetic background genome: built on the device (nrp_bg_build) when a context is given,
    and by the storage layer's host code for comparison."""
    if not cfg['genome']:
        return None, 0
    import cases
    from nrpcalc_b200 import encoding
    genome = cases.random_genome(cfg['genome'], cfg['seed'] * 11)
    buf, off = encoding.join_parts([genome])
    K = cfg['Lmax'] + 1
    if ctx is None:
        return np.unique(encoding.canonical_kmers(buf, off, K)), K
    ctx.bg_build(buf, off, K)                                    # warm-up (allocations)
    t0 = time.perf_counter()
    keys = ctx.bg_build(buf, off, K)
    t1 = time.perf_counter()
    host = np.unique(encoding.canonical_kmers(buf, off, K))
    t2 = time.perf_counter()
    assert keys.tolist() == host.tolist()
    if timing is not None:
        timing.update({'background_build_device_ms': 1e3 * (t1 - t0), 'background_build_host_numpy_ms': 1e3 * (t2 - t1),
                       'background_genome_bp': cfg['genome']})
    return keys, K


def algorithmic_bytes(total_len, m, key_bytes, packed):
    """Bytes each bulk kernel has to move per launch: a record is one 8-byte word when the k-mer and
    the value fit 64 bits (packed), else key + 4-byte value."""
    rec = 8 if packed else key_bytes + 4
    return {'histogram': total_len, 'split1': total_len + m * rec, 'split2': 2 * m * rec, 'filter': m * rec,
            'flags': total_len}


def run_reference_arm(args, cfg):
    """The reference's own graph-build path (nrpcalc/base/hgraph.py:202-249, get_homology_graph) on a bounded sample of the
    workload, one core (the reference is single-threaded pure Python).  The unmodified reference is imported from the offline
    install under baseline/_ref (or /root/reference in the development container) with Bio / RNA / ShareDB stubbed
    (tests/golden/refshim.py); without a reference tree the literal port oracle/ref_port.py is timed instead and the line says so."""
    import tempfile
    import cases
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import refshim
    k = cfg['Lmax'] + 1
    sample_parts = max(1000, min(cfg['n'], args.reference_parts))
    buf, off = make_workload(cfg, sample_parts)
    parts = cases.buffer_to_list(buf, off)
    m = int(np.maximum(np.diff(off) - k + 1, 0).sum())
    genome = cases.random_genome(min(cfg['genome'], 200_000), cfg['seed'] * 11) if cfg['genome'] else None
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix='nrp_refarm_')
    os.chdir(scratch)
    try:
        if refshim.reference_available():
            ref = refshim.load_reference()
            from nrpcalc.base import hgraph as ref_hgraph
            kind, where = 'reference', 'unmodified nrpcalc {} from {}'.format(getattr(ref, '__version__', '1.7.6'), os.path.relpath(refshim.REFERENCE_ROOT, ROOT)
                                                                                 if refshim.REFERENCE_ROOT.startswith(ROOT) else refshim.REFERENCE_ROOT)
            bg = None
            if genome:
                bg = ref.background(path=os.path.join(scratch, 'bg'), Lmax=cfg['Lmax'], verbose=False)
                bg.add(genome)

            def one_step():
                return ref_hgraph.get_homology_graph(list(parts), bg, k, cfg['internal_repeats'], os.path.join(scratch, 'seq_list.txt'), False)
        else:
            from oracle import ref_port
            kind, where = 'port', 'oracle/ref_port.py: literal pure-Python restatement (no reference tree on this box)'
            bg = None
            if genome:
                bg = ref_port.SetBackground(k)
                bg.multiadd([genome])

            def one_step():
                return ref_port.homology_graph(list(parts), bg, k, cfg['internal_repeats'])[0]
        times = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            one_step()
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    finally:
        os.chdir(cwd)
        import shutil
        shutil.rmtree(scratch, ignore_errors=True)
    ms = 1e3 * sum(times) / len(times)
    value = m / (ms * 1e-3)
    total_m = cfg['n'] * (m / sample_parts)
    sample = ('{} of {} parts ({} k-mers) per step through hgraph.get_homology_graph (table + clique tail + nx.Graph); the path is linear in '
              'the number of k-mers, so the full workload ({:.3g} k-mers) extrapolates to {:.0f} s per step at this rate').format(
                  sample_parts, cfg['n'], m, total_m, total_m / value)
    line = {'impl': 'reference', 'metric': 'finder_graph_build_kmers_per_s', 'value': value, 'unit': 'k-mers/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u64' if k > 16 else 'u32', 'data': 'synthetic',
            'config': {'workload': cfg['label'], 'n_parts': cfg['n'], 'Lmax': cfg['Lmax'], 'sample_parts': sample_parts,
                       'background_bp': len(genome) if genome else 0},
            'parts_per_s': sample_parts / (ms * 1e-3),
            'cpu_baseline': {'value': value, 'unit': 'k-mers/s', 'cores': 1, 'kind': kind, 'sample': sample, 'source': where,
                             'host_cores_available': os.cpu_count()},
            'e2e': {'value': value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def cpu_baseline(cfg, buf, off, k):
    """oracle/nrp_oracle.c on a bounded sample of the workload, all host cores."""
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    n_sample = min(cfg['n'], 400_000)
    sub_off = off[:n_sample + 1]
    sub_buf = buf[:int(sub_off[-1])]
    t0 = time.perf_counter()
    res = c_oracle.build_table(sub_buf, sub_off, k, cfg['internal_repeats'], None, 0, n_threads=cores, want_groups=True)
    dt = time.perf_counter() - t0
    return {'value': res['n_records'] / dt, 'unit': 'k-mers/s', 'cores': cores, 'kind': 'port',
            'sample': 'first {} of {} parts ({} k-mers), oracle/nrp_oracle.c hash-table port, {:.1f} s'.format(
                n_sample, cfg['n'], res['n_records'], dt)}


def verify_against_oracle(cfg_name, cfg, buf, off, ids, k, res, records, bg_keys, bg_K):
    """Full-size parity: the result the device produced for the WHOLE batch against oracle/nrp_oracle.c (checker only)."""
    from oracle import c_oracle
    import parity
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    want = c_oracle.build_table(buf, off, k, cfg['internal_repeats'], bg_keys, bg_K, n_threads=cores, want_groups=True,
                                n_slices=VERIFY_SLICES.get(cfg_name, 4))
    t1 = time.perf_counter()
    got = parity.canonical(res.flags, res.indptr, res.members, res.first_window, res.keys, res.last_single, res.edges, None, records)
    ref = parity.from_c_oracle(want, ids)
    cmp = parity.compare(got, ref)
    return {'vs_oracle': cmp['all'], 'components': cmp, 'digest': parity.digest(got), 'oracle_digest': parity.digest(ref),
            'checked': parity.summary(got), 'oracle': 'oracle/nrp_oracle.c on the full batch ({} parts), {} threads, {} key-space passes, {:.1f} s'.format(
                cfg['n'], cores, VERIFY_SLICES.get(cfg_name, 4), t1 - t0),
            'compare_s': time.perf_counter() - t1}


def finder_api_timing(n_parts, length, Lmax, seed):
    """What a user of the reference sees: nrpcalc_b200.finder(list[str], Lmax) wall clock on a cut of the workload, with the
    library's own split into device call, order-exact replay, recovery and marshalling (hgraph.LAST_TIMINGS)."""
    import random
    import tempfile
    import cases
    import nrpcalc_b200
    from nrpcalc_b200 import hgraph
    parts = cases.random_parts(n_parts, length, seed)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix='nrp_api_'))
    try:
        random.seed(1)
        t0 = time.perf_counter()
        found = nrpcalc_b200.finder(parts, Lmax, internal_repeats=False, vercov='nrp2', verbose=False)
        wall = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    info = dict(hgraph.last_build_info)
    run = dict(nrpcalc_b200._finder.last_run_info)
    # the same job through the packed entry (bytes in, uniquify on the device, no side file): same toolbox, less host work
    packed_block = None
    try:
        from nrpcalc_b200 import packed
        blob = ''.join(parts).encode('ascii')
        os.chdir(tempfile.mkdtemp(prefix='nrp_api_'))
        random.seed(1)
        t0 = time.perf_counter()
        found_packed = nrpcalc_b200.finder_packed(blob, Lmax, part_len=length, internal_repeats=False, vercov='nrp2')
        wall_packed = time.perf_counter() - t0
        packed_block = {'wall_s': round(wall_packed, 4), 'same_toolbox': found_packed == found and list(found_packed) == list(found)}
        packed_block.update({k: (round(v, 4) if isinstance(v, float) else v) for k, v in packed.last_run_info.items()})
    except Exception as exc:                                   # reported, never fatal for the bench line
        packed_block = {'error': repr(exc)}
    finally:
        os.chdir(cwd)
    return {'parts': n_parts, 'part_len': length, 'Lmax': Lmax, 'vercov': 'nrp2', 'wall_s': round(wall, 4), 'toolbox': len(found),
            'finder_packed': packed_block,
            'kmers': info.get('records'), 'device_ms': info.get('device_ms'),
            'marshal_and_device_call_s': round(info.get('call_s', 0.0), 4), 'clique_replay_and_graph_s': round(info.get('replay_s', 0.0), 4),
            'recovery_s': round(run.get('recovery_s', 0.0), 4), 'checks_sidefiles_output_s': round(max(0.0, wall - run.get('graph_s', 0.0) - run.get('recovery_s', 0.0)), 4),
            'native_tail': info.get('native_tail'),
            'note': 'host wall clock of the unchanged (order-exact) tail dominates; the graph-build metric is the device / C-ABI figure'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default=None)
    ap.add_argument('--reference-parts', type=int, default=20_000)
    ap.add_argument('--n-parts', type=int, default=0, help='run a SCALED copy of the configuration with this many parts (labelled as such; development runs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-edges', action='store_true')
    ap.add_argument('--verify', action='store_true', help='compare the full result with oracle/nrp_oracle.c (default for c1/c2 at N=1; N>1: also against the oracle, besides the single-GPU result)')
    ap.add_argument('--no-verify', action='store_true')
    ap.add_argument('--verify-parts', type=int, default=5_000_000, help='N > 1, batches beyond one device: parts of the cut the parity check runs on')
    ap.add_argument('--api-parts', type=int, default=100_000, help='parts of the finder() API timing leg (0 = skip)')
    ap.add_argument('--h2d-chunks', type=int, default=8, help='e2e leg: pieces of the streamed host-to-device copy (1 = one blocking copy)')
    ap.add_argument('--e2e-entry', choices=('auto', 'offsets', 'uniform'), default='auto',
                    help='e2e leg: nrp_load_parts[_streamed] with an offsets array, or nrp_load_parts_uniform (parts of one length; auto = uniform when the config has one length)')
    ap.add_argument('--opt', action='append', default=[], help='library tuning knob name=value (nrp_ctx_set_option), repeatable')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    multi = args.gpus > 1 or world > 1
    # N = 1: BASELINE config 2 (the configuration the metric is quoted on).  N > 1: BASELINE config 4, the 10M-part batch
    # the reference plan shards across 2/4/8 GPUs (strong scaling); --config c2w = one config-2 batch per GPU (weak scaling)
    cfg_name = args.config or ('c4' if multi else 'c2')
    args.weak = cfg_name == 'c2w'
    cfg = dict(CONFIGS['c2' if args.weak else cfg_name])
    if args.n_parts > 0 and args.n_parts != cfg['n']:
        cfg['label'] += ' -- SCALED to {} of {} parts'.format(args.n_parts, cfg['n'])
        cfg['n'] = args.n_parts

    if args.impl == 'reference':
        if rank == 0:
            run_reference_arm(args, cfg)
        return

    if multi:
        import bench_sharded
        bench_sharded.main(args, cfg, rank, world, local_rank)
        return

    import torch
    from nrpcalc_b200 import _native

    k = cfg['Lmax'] + 1
    buf, off = make_workload(cfg)
    ids = np.arange(cfg['n'], dtype=np.uint32)[::-1].copy()      # processing order of a duplicate-free list: ids N-1..0
    want_edges = not args.no_edges

    torch.cuda.set_device(local_rank)
    ctx = _native.Context(local_rank)
    for item in args.opt:
        name, value = item.split('=')
        ctx.set_option(name, int(value))
    bg_timing = {}
    bg_keys, bg_K = background_keys(cfg, ctx, bg_timing)
    ctx.set_background(bg_keys, bg_K)
    # pinned host staging for the e2e leg; device-resident copy for the kernel-only leg
    host_ascii = torch.from_numpy(buf.copy()).pin_memory()
    dev_ascii = host_ascii.cuda(non_blocking=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')     # > 126 MB L2

    # ---- device-resident leg
    ctx.load_parts(None, off, ids, device_ptr=dev_ascii.data_ptr())
    sampler = ClockSampler(local_rank)
    sampler.start()                                  # nvidia-smi needs a moment before its first row: started before the warm-up
    for _ in range(args.warmup):
        ctx.build(k, cfg['internal_repeats'], want_edges)
    torch.cuda.synchronize()
    sampler.mark()
    stage_ms = {}
    total_ms = 0.0
    launches = 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.build(k, cfg['internal_repeats'], want_edges)
        t = ctx.timings()
        total_ms += t['total']
        for name, v in t.items():
            stage_ms[name] = stage_ms.get(name, 0.0) + v
        launches += ctx.counts()['launches']
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    sampler.top_up(lambda: ctx.build(k, cfg['internal_repeats'], want_edges))
    clocks = sampler.stop()                          # the clocks belong to the device-timed region; no polling during the e2e leg
    counts = ctx.counts()
    m = counts['records']
    ms_per_step = total_ms / args.steps
    value = m / (ms_per_step * 1e-3)

    # ---- end-to-end leg: host buffers in, host results out
    e2e_times = []
    e2e_parts = []
    h2d = d2h = 0
    uniform_entry = args.e2e_entry == 'uniform' or (args.e2e_entry == 'auto' and not isinstance(cfg['length'], tuple))
    host_ids = torch.from_numpy(ids.copy()).pin_memory()
    for step in range(2 + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if uniform_entry:
            ctx.load_parts_uniform(host_ascii.numpy(), cfg['n'], cfg['length'], host_ids.numpy(), chunks=args.h2d_chunks)
        else:
            ctx.load_parts(host_ascii.numpy(), off, ids, chunks=args.h2d_chunks)
        t1 = time.perf_counter()
        ctx.build(k, cfg['internal_repeats'], want_edges, early_fetch=True)       # as nrpcalc_b200.hgraph calls it: the result is fetched next
        t2 = time.perf_counter()
        res = ctx.fetch(want_keys=True)
        dt = time.perf_counter() - t0
        if step >= 2:
            e2e_times.append(dt)
            e2e_parts.append((t1 - t0, t2 - t1, t0 + dt - t2))
        # parts of one length: the offsets are generated on the device, only the bytes and the ids travel
        h2d = buf.nbytes + ids.nbytes + (0 if not isinstance(cfg['length'], tuple) else off.nbytes)
        d2h = (res.flags.nbytes + res.indptr.nbytes + res.members.nbytes + res.first_window.nbytes + res.keys.nbytes +
               res.last_single.nbytes + (res.edges[0].nbytes + res.edges[1].nbytes if res.edges else 0) + 16 * 8 * 2)
    e2e_value = m / (sum(e2e_times) / len(e2e_times))

    # ---- parity of the FULL result against the C oracle (checker only; default for the configs it does in seconds)
    parity_block = None
    do_verify = (args.verify or cfg_name in ('c1', 'c2')) and not args.no_verify
    if do_verify:
        ctx.load_parts(None, off, ids, device_ptr=dev_ascii.data_ptr())
        ctx.build(k, cfg['internal_repeats'], True)
        full = ctx.fetch(want_keys=True, copy=True)
        parity_block = verify_against_oracle(cfg_name, cfg, buf, off, ids, k, full, ctx.counts()['records'], bg_keys, bg_K)
        del full
    api_block = None
    if args.api_parts > 0 and not isinstance(cfg['length'], tuple):
        ctx.set_background(None, 0)
        api_block = finder_api_timing(min(args.api_parts, cfg['n']), cfg['length'], cfg['Lmax'], cfg['seed'])

    # ---- roofline of the dominant kernel
    peak, peak_src = measured_peak_gbs()
    algo = algorithmic_bytes(int(off[-1]), m, counts['key_bytes'], counts['packed'])
    stage_avg = {name: stage_ms.get(name, 0.0) / args.steps for name in algo}
    dominant = max(stage_avg, key=lambda name: stage_avg[name])
    kernel_of = {'histogram': 'k_hist_from_ascii', 'split1': 'k_split_uniform' if counts.get('uniform_kernel') else 'k_split_from_ascii', 'split2': 'k_split_from_records', 'filter': {0: 'k_filter_bloom', 1: 'k_filter_warp', 2: 'k_filter_warp', 3: 'k_filter_warp'}.get(counts.get('filter_warp', 0), 'k_filter_bloom'),
                 'flags': 'k_flag_windows'}
    achieved = algo[dominant] / (stage_avg[dominant] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(cfg_name, {}).get(kernel_of[dominant])
    # stages that did not run (no histogram pass on the direct path, no flag pass for odd k without background) move no bytes
    pipeline_bytes = sum(algo[name] for name in ('histogram', 'split1', 'split2', 'filter', 'flags') if stage_avg[name] > 0)
    roofline = {'bound': 'hbm', 'kernel': kernel_of[dominant], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': algo[dominant],
                'kernel_ms': stage_avg[dominant],
                'per_stage': {name: {'ms': stage_avg[name], 'GB/s': (algo[name] / (stage_avg[name] * 1e-3) / 1e9) if stage_avg[name] > 0 else None}
                              for name in algo},
                'pipeline': {'bytes_per_kmer': pipeline_bytes / max(m, 1), 'GB/s': pipeline_bytes / (ms_per_step * 1e-3) / 1e9,
                             'frac': pipeline_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                             'survey_formula_bytes_per_kmer': (int(off[-1]) / max(m, 1)) + (counts['key_bytes'] + 4) * (2 * -(-2 * k // 8) + 3) + counts['key_bytes'],
                             }}
    roofline['pipeline']['survey_formula_frac'] = roofline['pipeline']['survey_formula_bytes_per_kmer'] * m / (ms_per_step * 1e-3) / 1e9 / peak

    line = {'metric': 'finder_graph_build_kmers_per_s', 'value': value, 'unit': 'k-mers/s', 'n_gpus': 1, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u64' if counts['key_bytes'] == 8 else 'u32', 'data': 'synthetic',
            'config': {'workload': cfg['label'], 'n_parts': cfg['n'], 'part_len': cfg['length'], 'Lmax': cfg['Lmax'], 'k': k,
                       'internal_repeats': cfg['internal_repeats'], 'background_kmers': 0 if bg_keys is None else int(bg_keys.size),
                       'kmers': m, 'edges': want_edges, 'record_bytes': 8 if counts['packed'] else counts['key_bytes'] + 4,
                       'l2': 'explicit 256 MiB flush between steps; per-step working set {:.2f} GB'.format(
                           2 * m * (8 if counts['packed'] else counts['key_bytes'] + 4) / 1e9)},
            'parts_per_s': cfg['n'] / (ms_per_step * 1e-3),
            'e2e': {'value': e2e_value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': 1e3 * sum(e2e_times) / len(e2e_times),
                    'ms_load_build_fetch': [1e3 * sum(p[i] for p in e2e_parts) / len(e2e_parts) for i in range(3)],
                    'h2d_chunks': args.h2d_chunks,
                    'path': 'C ABI, page-locked host buffers: ' + ('nrp_load_parts_uniform (no offsets array; {} piece(s), copies only enqueued when > 1)'.format(args.h2d_chunks)
                                                                  if uniform_entry else ('nrp_load_parts_streamed' if args.h2d_chunks > 1 else 'nrp_load_parts')) +
                            ' + nrp_build (early_fetch: results that are final after the grouping stage travel beside the rest of the build) + nrp_result_host'},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline,
            'stage_ms': {name: v / args.steps for name, v in stage_ms.items()},
            'result': {name: counts[name] for name in ('records', 'excluded', 'groups', 'members', 'edges', 'candidates', 'runs', 'buckets', 'oversized',
                                                       'direct', 'fallbacks', 'packed', 'uniform_kernel')},
            'wall_s_timed_region': wall}
    if bg_timing:
        line['config'].update(bg_timing)
    if parity_block is not None:
        line['parity'] = parity_block
    if api_block is not None:
        line['finder_api'] = api_block
    if not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(cfg, buf, off, k)
    print(json.dumps(line))
    if parity_block is not None and not parity_block['vs_oracle']:
        sys.stderr.write('PARITY MISMATCH against oracle/nrp_oracle.c: {}\n'.format(parity_block['components']))
        sys.exit(3)


if __name__ == '__main__':
    main()
