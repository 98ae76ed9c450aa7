#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (genome-pairs/s, fragments mapped/s, % of HBM roofline) on
BASELINE.json configs[1]: synthetic 5 Mbp query genomes against an index of 1,000 synthetic 5 Mbp
references mutated to 80-99 % identity, on N B200s.

The job is the same for every N (strong scaling): a FIXED list of Q = 64 distinct query genomes --
the base genome and 63 independent light mutations of it, so every (query, reference) pair stays in
the 80-99 % band -- is mapped against the 1,000-reference index.  The list is dealt to the N ranks
with `sharding.partition_queries` (longest-processing-time by fragment count, replicated index, no
collective in the mapping path); N = 1 maps the whole list on one GPU.

A step = one pass of the hot path over the whole list:  Mapper.query_many(this rank's share)  =
Q x 1000 genome pairs, Q x 1,666 fragments.  The index build (sketch all references + index) is
set-up, as in the reference's own benchmark (benches/mapping/bench.py:34-66), and is reported beside
the metric with its own roofline.

  python bench.py [--gpus N --steps K --warmup W]           our arm (one JSON line)
  python bench.py --impl reference [...]                    the reference's CPU code on the host cores
  python bench.py --shard references [...]                  N > 1: the configs[4] layout on this workload

`value` times K steps with the queries already resident in HBM: one CUDA-event pair around every
step on the library's stream, max over ranks.  `e2e` times the same K steps through pyfastani_b200's
public API with HOST buffers (pinned staging, H2D and D2H inside the timed region, wall clock between
barriers, max over ranks).
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC, UNIT = "genome_pairs_per_s", "genome-pairs/s"
FRAG = 3000
QUERY_IDENTITY = 0.998          # queries 1..Q-1 are the base genome mutated to this identity


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout: whatever libraries print there (the NCCL version banner, ...)
    goes to stderr instead; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refs", type=int, default=1000)
    ap.add_argument("--queries", type=int, default=64, help="Q: distinct query genomes in the fixed list (the same for every N)")
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--cpu-sample-refs", type=int, default=16)
    ap.add_argument("--cpu-linearity", default="16,32,64",
                    help="--impl reference: sampled-reference counts whose pairs/s must agree (the sample extrapolates linearly)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="queries", choices=["queries", "references"],
                    help="N > 1: 'queries' = replicated index, the query list partitioned over the GPUs (the default); "
                         "'references' = whole reference genomes split over the GPUs, every GPU maps every query against its shard, "
                         "hit rows all-gathered over NCCL inside the library and merged (SURVEY.md 8(e), BASELINE configs[4] layout)")
    ap.add_argument("--profile-build", action="store_true",
                    help="bracket Sketch.index() with cudaProfilerStart/Stop, print its timers and exit (launch list of the index build)")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the device-timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args(argv)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(kernel, a):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` capture
    of this workload (profiles/traffic.json, written by profiles/make_traffic.py); None for any other workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t.get("refs") == a.refs and t.get("length") == a.length:
            return t["kernels"].get(kernel)
    except Exception:
        pass
    return None


def workload_name(a):
    return ("configs[1] 1-to-many, %d distinct queries per step: synthetic %.1f Mbp query genomes (the base genome and %d mutations of it "
            "at %.1f %%) vs %d synthetic %.1f Mbp references at 80-99%% identity"
            % (a.queries, a.length / 1e6, a.queries - 1, 100 * QUERY_IDENTITY, a.refs, a.length / 1e6))


# ---------------------------------------------------------------------------------------------
# torch.distributed plumbing (NCCL on the GPU box, gloo in tests/test_sharding.py)
# ---------------------------------------------------------------------------------------------
def max_over_ranks(dist, torch, dev, x, world):
    if world == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, torch, dev, x, world):
    if world == 1:
        return int(x)
    t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    return int(t.item())


def broadcast_bytes(dist, torch, dev, data, n, world):
    """`data` (bytes, rank 0) -> the same n bytes on every rank."""
    if world == 1:
        return data
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if data is not None:
        t.copy_(torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()))
    dist.broadcast(t, src=0)
    return t.cpu().numpy().tobytes()


# ---------------------------------------------------------------------------------------------
# synthetic data: the generator of tests/synth.py, with the per-reference
# mutation drawn on the GPU so that setting up 5 Gbp of references takes seconds, not minutes
# ---------------------------------------------------------------------------------------------
def base_codes(a):
    return np.random.default_rng(a.seed).integers(0, 4, size=a.length, dtype=np.uint8)


def identities(a):
    return np.linspace(0.80, 0.99, a.refs)


def query_bytes(a, base, q):
    """Query q of the fixed list: the base genome itself (q = 0) or an independent mutation of it."""
    import synth
    if q == 0:
        return synth.to_bytes(base)
    rng = np.random.default_rng(a.seed * 7_000_003 + q)
    return synth.to_bytes(synth.mutate_codes(rng, base, QUERY_IDENTITY))


def reference_on_device(torch, base_dev, lut, ident, seed, index):
    g = torch.Generator(device=base_dev.device)
    g.manual_seed(seed * 1_000_003 + index)
    hit = torch.rand(base_dev.shape, generator=g, device=base_dev.device) < (1.0 - ident)
    shift = torch.randint(1, 4, base_dev.shape, generator=g, device=base_dev.device, dtype=torch.uint8)
    codes = (base_dev + hit.to(torch.uint8) * shift) & 3
    return lut[codes.long()]


def reference_on_host(a, base, index):
    import synth
    rng = np.random.default_rng(a.seed * 1_000_003 + index)
    return synth.to_bytes(synth.mutate_codes(rng, base, float(identities(a)[index])))


KERNEL_NAMES = {"ms_sketch": "sketch_kernel", "ms_lookup": "lookup_kernel", "ms_seed_sort": "fill_seeds+DeviceRadixSort",
                "ms_l1": "l1_fused_kernel", "ms_l2_prep": "l2_prep_kernel", "ms_l2_events": "l2_events_kernel",
                "ms_l2_slide": "l2_slide_kernel", "ms_cgi": "cgi_best_kernel"}


def stage_bytes(inf, bases):
    """Algorithmic bytes per stage from the realised counters of `inf`
    (Mapper.last_query_info, possibly summed over calls) and the query bases they cover."""
    s_mean = inf["sketch_sum"] / max(inf["fragments"], 1)
    return {
        "ms_sketch": 1.96 * bases,
        "ms_lookup": 36.0 * inf["sketch_sum"],
        # L1 on chip (l1_fused_kernel): the position lists once (4 B / seed), one 4-byte gpos gather per seed, 16 B per
        # candidate; ms_seed_sort is the device-wide sort of the fragments that do not fit on chip (none here)
        "ms_seed_sort": 16.0 * inf["seeds"] * inf["l1_sorted_fragments"] / max(inf["fragments"], 1),
        "ms_l1": 8.0 * inf["seeds"] + 16.0 * inf["candidates"],
        # L2 = prep (index searches) + events (classify: reads the (hash, order word) stream once, writes 2-byte events
        # and the start state) + slide (replays the events from the start window until both sides are pruned)
        "ms_l2_prep": 56.0 * inf["candidates"],
        "ms_l2_events": 8.0 * inf["scanned"] + 2.0 * inf["events"] + (4.0 * s_mean + 16.0) * inf["candidates"],
        "ms_l2_slide": 2.0 * inf["events_replayed"] + (s_mean + 54.0) * inf["candidates"],
        "ms_cgi": 16.0 * inf["candidates"],
    }


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref) or, if absent, the C port
# ---------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's CPU implementation of the path on the host cores: an index over a SAMPLE of the references
    (evenly spaced identities), mapped with whole queries of the list.  BASELINE.md section 3 asks for both ways of using
    the cores -- the reference's own intra-query worker pool (`threads = cores`) and a thread pool over queries with one
    thread each -- and for the single-thread figure; the faster schedule is the arm's step."""

    def __init__(self, sample_refs):
        from oracle.oracle import Oracle, available
        self.kind = "reference" if "reference" in available() else "port"
        self.orc = Oracle(self.kind)
        self.cores = os.cpu_count() or 1
        self.ids = [i for i, _ in sample_refs]
        t0 = time.perf_counter()
        self.sk = self.orc.sketch()
        # sketching the sample is set-up: the reference's addMinimizers per genome, one genome per host thread
        self.sk.add_genomes(self.ids, [r for _, r in sample_refs], threads=self.cores)
        self.sk.index()
        self.t_index = time.perf_counter() - t0

    def intra(self, query, threads=None):
        kw = {"threads": threads or self.cores} if self.kind == "reference" else {}
        return self.sk.query_genome(query, **kw)[0]

    def pool(self, queries):
        kw = {"threads": 1} if self.kind == "reference" else {}
        with ThreadPoolExecutor(max_workers=min(self.cores, len(queries))) as ex:
            return list(ex.map(lambda q: self.sk.query_genome(q, **kw)[0], queries))

    def time_schedules(self, queries, with_t1=True):
        """pairs/s of the three schedules on `queries` (full genomes) against the sampled index."""
        n = len(self.ids)
        out = {}
        t0 = time.perf_counter(); self.intra(queries[0]); out["intra_query_pool_T%d" % self.cores] = n / (time.perf_counter() - t0)
        if self.kind == "reference":
            qs = [queries[i % len(queries)] for i in range(min(self.cores, max(len(queries), 1) * 4, 32))]
            t0 = time.perf_counter(); self.pool(qs); out["thread_pool_over_queries_T1_each"] = n * len(qs) / (time.perf_counter() - t0)
            if with_t1:
                t0 = time.perf_counter(); self.intra(queries[0], threads=1); out["T1"] = n / (time.perf_counter() - t0)
        return out


def sample_ids(a, n=None):
    n = max(1, min(n or a.cpu_sample_refs, a.refs))
    return sorted({int(round(x)) for x in np.linspace(0, a.refs - 1, n)})


def config1_case():
    """BASELINE configs[0]: the reference's own runnable case, E. coli K12 MG1655 (query) vs the Shigella flexneri 2a
    draft (reference), from the fixtures under tests/golden/data (the vendored FastANI genomes, gzipped)."""
    d = os.path.join(ROOT, "tests", "golden", "data")
    try:
        ecoli = gzip.open(os.path.join(d, "ecoli.seq.gz")).read().split(b"\n")
        shig = gzip.open(os.path.join(d, "shigella.seq.gz")).read().split(b"\n")
        return [x for x in ecoli if x], [x for x in shig if x]
    except Exception:
        return None, None


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the same path on the host cores.  A step = whole query
    genomes of the list against a bounded SAMPLE of the references, with the schedule that uses the cores best; K timed
    steps after W warm-up steps, exactly as the other arm counts them."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = base_codes(a)
    nq = max(1, min(a.queries, 4))
    queries = [query_bytes(a, base, q) for q in range(nq)]
    lin_counts = sorted({int(x) for x in a.cpu_linearity.split(",") if x.strip()} | {a.cpu_sample_refs})
    arms = {}
    for n in lin_counts:
        refs = [(i, reference_on_host(a, base, i)) for i in sample_ids(a, n)]
        arms[n] = CpuArm(refs)
        del refs
    arm = arms[a.cpu_sample_refs]
    sched = arm.time_schedules(queries)
    linearity = {str(n): arms[n].time_schedules(queries, with_t1=False) for n in lin_counts}
    best = max((k for k in sched if k != "T1"), key=lambda k: sched[k])
    use_pool = best.startswith("thread_pool")
    step_queries = [queries[i % nq] for i in range(min(arm.cores, 32))] if use_pool else [queries[0]]
    pairs = len(arm.ids) * len(step_queries)

    def step(i):
        if use_pool:
            arm.pool(step_queries)
        else:
            arm.intra(queries[i % nq])

    for i in range(a.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(i)
    dt = (time.perf_counter() - t0) / max(a.steps, 1)
    value = pairs / dt
    lin_vals = [linearity[str(n)][best] for n in lin_counts]
    # configs[0], in full: the real E. coli / Shigella pair
    c1 = None
    ecoli, shig = config1_case()
    if ecoli and arm.kind == "reference":
        sk = arm.orc.sketch()
        t0 = time.perf_counter(); sk.add_draft("shigella", shig); sk.index(); t_ix = time.perf_counter() - t0
        t0 = time.perf_counter(); h, _ = sk.query_draft(ecoli, threads=arm.cores); t_q = time.perf_counter() - t0
        t0 = time.perf_counter(); sk.query_draft(ecoli, threads=1); t_q1 = time.perf_counter() - t0
        c1 = {"workload": "configs[0] E. coli K12 MG1655 (query) vs Shigella flexneri 2a draft (reference), in full",
              "index_s": t_ix, "query_s": t_q, "query_s_T1": t_q1, "pairs_per_s": 1.0 / t_q,
              "hit": [int(h[0]["matches"]), int(h[0]["fragments"]), float(h[0]["identity"])] if len(h) else None}
    cb = {"value": value, "unit": UNIT, "cores": arm.cores if arm.kind == "reference" else 1, "kind": arm.kind,
          "sample": ("%s: %d whole %.1f Mbp quer%s of the list vs %d of the %d references (evenly spaced identities) per step, index prebuilt "
                     "(%.1f s), %.2f s per step; pairs/s at %s sampled references: %s (max/min %.3f: the cost is linear in related references)"
                     % (best, len(step_queries), a.length / 1e6, "ies" if len(step_queries) > 1 else "y", len(arm.ids), a.refs, arm.t_index, dt,
                        "/".join(str(n) for n in lin_counts), "/".join("%.1f" % v for v in lin_vals), max(lin_vals) / max(min(lin_vals), 1e-9))),
          "schedules": sched, "linearity": linearity}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "refs": a.refs, "queries": a.queries, "length": a.length, "fragment_length": FRAG, "k": 16},
            "cpu_baseline": cb,
            "fragments_per_s": len(step_queries) * (a.length // FRAG) / dt,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "config1": c1}
    emit(line)


# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    import pyfastani_b200 as pf
    import synth
    from pyfastani_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- set-up: references generated in HBM, sketched and indexed (not in the timed region) ----
    base = base_codes(a)
    base_dev = torch.from_numpy(base).to(dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    idents = identities(a)
    want_cpu = rank == 0 and not a.no_cpu_baseline
    keep = set(sample_ids(a)) if want_cpu else set()
    sample_refs = []
    sketch = pf.Sketch(device=local)
    by_refs = a.shard == "references" and world > 1
    offsets = sharding.reference_shards([a.length] * a.refs, world) if by_refs else [0, a.refs]
    my_refs = range(offsets[rank], offsets[rank + 1]) if by_refs else range(a.refs)
    comm = None
    if by_refs:
        # the library's own NCCL communicator; its id travels through the torch group that the contract's barrier uses
        comm = sharding.connect(world, rank, local, exchange=lambda uid: broadcast_bytes(dist, torch, dev, uid, 128, world))
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    batch, names = [], []
    for i in my_refs:
        ref = reference_on_device(torch, base_dev, lut, float(idents[i]), a.seed, i)
        if i in keep:
            sample_refs.append((i, ref.cpu().numpy().tobytes()))
        batch.append(ref); names.append(i)
        if len(batch) == 16 or i == my_refs[-1]:
            torch.cuda.synchronize(dev)
            sketch.add_many(names, [pf.DeviceSequence.from_pointer(r.data_ptr(), r.numel(), local, r) for r in batch])
            batch, names = [], []
    t_sketch = time.perf_counter() - t0
    sk_stats = sketch.build_stats
    n_min = len(sketch.minimizers)
    if a.profile_build:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    mapper = sketch.index()
    t_index = time.perf_counter() - t0
    if a.profile_build:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        emit({"index_build": {"wall_index_s": t_index, "minimizers": n_min, "stats": mapper.build_stats}})
        return
    ix_stats = mapper.build_stats
    del base_dev
    torch.cuda.empty_cache()

    # ---- the fixed query list and this rank's share ---------------------------------------------
    Q = a.queries
    frags_per_query = a.length // FRAG
    if by_refs:
        mine = list(range(Q))                         # every rank maps every query against its shard
    else:
        mine = sharding.partition_queries([frags_per_query] * Q, world)[rank]
    host_q = [query_bytes(a, base, q) for q in mine]
    dev_q = [pf.DeviceSequence.from_host(b, local) for b in host_q]

    def step(items):
        """One step on this rank: its share of the list through ONE library call (fa_query_batch: the next query is
        staged while the current one is mapped; reference-sharded: fa_query_batch_sharded, rows gathered over NCCL)."""
        if by_refs:
            return mapper.query_many(items, rows=True, comm=comm, genome_offsets=offsets)
        return mapper.query_many(items, rows=True) if items else []

    hits = None
    for _ in range(max(a.warmup, 1)):
        hits = step(dev_q)
        step(host_q)

    # ---- timed: K steps, queries resident in HBM (one CUDA-event pair per step inside the library) ----
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    stage, counters = {}, {}
    dev_ms, launches = 0.0, 0
    if a.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
    t_wall = time.perf_counter()
    for _ in range(a.steps):
        step(dev_q)
        inf = mapper.last_query_info if mine else {}
        dev_ms += inf.get("ms_batch", 0.0)
        launches += inf.get("kernel_launches", 0)
        for k, v in inf.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    if a.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    if by_refs:     # the gather is part of the step: wall clock between the barriers instead of the library's event pairs
        dev_ms = wall_ms
    dev_ms = max_over_ranks(dist, torch, dev, dev_ms, world)
    wall_ms = max_over_ranks(dist, torch, dev, wall_ms, world)
    launches = sum_over_ranks(dist, torch, dev, launches, world)
    inf = dict(mapper.last_query_info) if mine else {}

    # ---- timed: the same K steps through the public API with host buffers (wall clock) ----------
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(a.steps):
        hits_e2e = step(host_q)
        if mine:
            h2d += mapper.last_query_info["h2d_bytes"]
            d2h += mapper.last_query_info["d2h_bytes"]
    barrier()
    e2e_s = max_over_ranks(dist, torch, dev, time.perf_counter() - t0, world)
    # ---- and once more with the queries packed to two bits per base on the host (pf.PackedSequence, packed outside the
    # timed region as a sequence store would hold them): a quarter of the bytes through host memory and PCIe ----------
    packed_q = [pf.PackedSequence.pack(b) for b in host_q]
    step(packed_q)
    barrier()
    t0 = time.perf_counter()
    h2d_pk = 0
    for _ in range(a.steps):
        hits_pk = step(packed_q)
        if mine:
            h2d_pk += mapper.last_query_info["h2d_bytes"]
    barrier()
    e2e_pk_s = max_over_ranks(dist, torch, dev, time.perf_counter() - t0, world)
    h2d_pk = sum_over_ranks(dist, torch, dev, h2d_pk, 1 if by_refs else world)
    assert len(hits_pk) == len(hits) and all(np.array_equal(x, y) for x, y in zip(hits_pk, hits))
    clocks = sampler.summary()
    h2d = sum_over_ranks(dist, torch, dev, h2d, 1 if by_refs else world)
    d2h = sum_over_ranks(dist, torch, dev, d2h, 1 if by_refs else world)
    assert len(hits_e2e) == len(hits) and all(np.array_equal(x, y) for x, y in zip(hits_e2e, hits))
    n_hits = sum(len(h) for h in hits)
    if not by_refs:
        n_hits = sum_over_ranks(dist, torch, dev, n_hits, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (SURVEY.md 8(d) algorithmic bytes, realised counters of one step of this
    # rank; a heavy query runs alone in its pass, so one launch = one query) ------------------------------------------
    peak, which = peaks()
    nq_rank = max(len(mine), 1)
    alg = stage_bytes(inf, a.length * nq_rank)
    per_step = {k: v / a.steps for k, v in stage.items()}
    top = max(alg, key=lambda k: per_step.get(k, 0.0))
    top_ms = per_step[top]
    achieved = alg[top] / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
    kernel_names = KERNEL_NAMES
    step_ms_rank = per_step.get("ms_batch", dev_ms / a.steps)
    roofline = {"bound": "hbm", "kernel": kernel_names[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(kernel_names[top], a),
                "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                "algorithmic_bytes_per_launch": alg[top] / nq_rank, "ms_per_launch": top_ms / nq_rank,
                "launches_per_step": nq_rank, "share_of_step": top_ms / step_ms_rank}
    stage_roofline = {k: {"ms_per_query": per_step.get(k, 0.0) / nq_rank, "alg_bytes_per_query": alg[k] / nq_rank,
                          "frac": (alg[k] / (per_step[k] * 1e-3) / 1e9 / peak) if per_step.get(k, 0) > 0 else None} for k in alg}
    whole = sum(alg.values())
    stage_roofline["whole_step"] = {"ms_per_query": step_ms_rank / nq_rank, "alg_bytes_per_query": whole / nq_rank,
                                    "frac": whole / (step_ms_rank * 1e-3) / 1e9 / peak}

    # ---- CPU baseline + full-size parity on the sampled pairs -----------------------------------
    cpu = parity = c1 = None
    if sample_refs and 0 in mine:
        arm = CpuArm(sample_refs)
        q0 = host_q[mine.index(0)]
        t0 = time.perf_counter(); ohits = arm.intra(q0); dt = time.perf_counter() - t0
        sched = {"intra_query_pool_T%d" % arm.cores: len(arm.ids) / dt}
        if arm.kind == "reference":
            qs = [host_q[i % len(host_q)] for i in range(min(arm.cores, 32))]
            t0 = time.perf_counter(); arm.pool(qs); sched["thread_pool_over_queries_T1_each"] = len(arm.ids) * len(qs) / (time.perf_counter() - t0)
        best = max(sched, key=lambda k: sched[k])
        cpu = {"value": sched[best], "unit": UNIT, "cores": arm.cores if arm.kind == "reference" else 1, "kind": arm.kind,
               "sample": "%s: whole %.1f Mbp queries of the list vs %d of the %d references (evenly spaced identities), index prebuilt (%.1f s)"
                         % (best, a.length / 1e6, len(arm.ids), a.refs, arm.t_index),
               "schedules": sched}
        rows0 = hits[mine.index(0)]
        mine_rows = {int(r["ref_genome"]): (int(r["matches"]), int(r["fragments"]), np.float32(r["identity"])) for r in rows0}
        ok = 0
        for h in ohits:
            gid = arm.ids[int(h["ref_genome"])]
            ok += int(mine_rows.get(gid) == (int(h["matches"]), int(h["fragments"]), np.float32(h["identity"])))
        parity = {"pairs_checked": len(ohits), "identical": ok,
                  "what": "query 0 of the list, hits of the sampled references: matches, fragments and identity bit-exact vs the CPU " + arm.kind}
        assert ok == len(ohits), "GPU hits differ from the CPU reference on the sampled pairs"
        # configs[0] in full on both arms: the real E. coli / Shigella pair
        ecoli, shig = config1_case()
        if ecoli and arm.kind == "reference":
            sk = pf.Sketch(device=local)
            t0 = time.perf_counter(); sk.add_draft("shigella", shig); m1 = sk.index(); t_ix = time.perf_counter() - t0
            m1.query_draft(ecoli)
            t0 = time.perf_counter(); h1 = m1.query_draft(ecoli); t_q = time.perf_counter() - t0
            osk = arm.orc.sketch()
            t0 = time.perf_counter(); osk.add_draft("shigella", shig); osk.index(); t_cix = time.perf_counter() - t0
            t0 = time.perf_counter(); oh, _ = osk.query_draft(ecoli, threads=arm.cores); t_cq = time.perf_counter() - t0
            same = len(h1) == len(oh) == 1 and (h1[0].matches, h1[0].fragments, np.float32(h1[0].identity)) == \
                (int(oh[0]["matches"]), int(oh[0]["fragments"]), np.float32(oh[0]["identity"]))
            c1 = {"workload": "configs[0] E. coli K12 MG1655 (query) vs Shigella flexneri 2a draft (reference), in full, host buffers",
                  "gpu": {"index_s": t_ix, "query_s": t_q, "ms_device": m1.last_query_info["ms_total"]},
                  "cpu_reference": {"index_s": t_cix, "query_s": t_cq, "cores": arm.cores},
                  "hit": [h1[0].matches, h1[0].fragments, h1[0].identity] if h1 else None, "identical": bool(same)}
            assert same, "configs[0]: GPU hit differs from the CPU reference"

    ms_per_step = dev_ms / a.steps
    pairs = Q * a.refs
    line = {
        "metric": METRIC, "value": pairs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 1), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "refs": a.refs, "queries": Q, "length": a.length, "fragment_length": FRAG, "k": 16,
                   "window": mapper.window_size,
                   "parallelism": ("reference genomes sharded over the GPUs (%s per rank), every GPU maps all %d queries against its shard, hit rows "
                                   "all-gathered over NCCL inside the library (fa_query_batch_sharded)" % ([offsets[r + 1] - offsets[r] for r in range(world)], Q)) if by_refs
                                  else "replicated index, the query list partitioned over the GPUs (sharding.partition_queries, %s queries per rank)"
                                       % [len(s) for s in sharding.partition_queries([frags_per_query] * Q, world)],
                   "l2_policy": "inputs larger than L2: the index is %.1f GB, every query streams it" % (n_min * 50 / 1e9)},
        "fragments_per_s": Q * frags_per_query / (ms_per_step * 1e-3),
        "wall_ms_per_step_resident": wall_ms / a.steps,
        "clocks": clocks,
        "e2e": {"value": pairs * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // a.steps,
                "d2h_bytes_per_step": d2h // a.steps, "ms_per_step": e2e_s / a.steps * 1e3,
                "fragments_per_s": Q * frags_per_query * a.steps / e2e_s},
        "e2e_packed": {"value": pairs * a.steps / e2e_pk_s, "unit": UNIT, "h2d_bytes_per_step": h2d_pk // a.steps,
                       "ms_per_step": e2e_pk_s / a.steps * 1e3,
                       "what": "the e2e steps with the host queries held at two bits per base (PackedSequence); hit rows identical"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stages": stage_roofline,
        "counters_rank0_per_step": {k: inf[k] for k in ("fragments", "sketch_sum", "seeds", "candidates", "scanned", "events", "mappings",
                                                        "l2_fallback", "l1_sorted_fragments", "events_replayed", "queries", "l1_parts")},
        "hits": n_hits,
        "parity": parity,
        "config1": c1,
        "index_build": {
            "wall_sketch_s": t_sketch, "wall_index_s": t_index, "minimizers": n_min,
            "sketch": {"ms": sk_stats["ms_sketch"], "bases": sk_stats["bases"], "alg_bytes": 1.96 * sk_stats["bases"],
                       "frac": 1.96 * sk_stats["bases"] / max(sk_stats["ms_sketch"] * 1e-3, 1e-12) / 1e9 / peak,
                       "mbp_per_s": sk_stats["bases"] / 1e6 / max(sk_stats["ms_sketch"] * 1e-3, 1e-12)},
            "index": {"ms": ix_stats["ms_build"], "ms_sort": ix_stats["ms_sort"], "alg_bytes": 28.0 * n_min,
                      "frac": 28.0 * n_min / max(ix_stats["ms_build"] * 1e-3, 1e-12) / 1e9 / peak},
            "what": "CUDA events on the library's stream: sketch = all launches of fa_sketch_add_genomes (1.96 B/base); "
                    "index = build_index (28 B/minimizer), ms_sort = its cub::DeviceRadixSort pair sort"},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
