#!/usr/bin/env python
"""Secondary measurement (the driver's contract line is bench.py): the many-to-many shapes of BASELINE configs[2]-[4] --
G synthetic genomes of 3-6 Mbp in a tree (bench_workloads.py), every genome mapped against the index of all of them.

    python bench_many.py [--genomes 256] [--tree genus|related] [--drafts] [--cpu-sample 2]          one GPU
    torchrun --nproc-per-node N bench_many.py --shard queries ...        configs[3]: replicated index, queries dealt to the GPUs
    torchrun --nproc-per-node N bench_many.py --shard references ...     configs[4]: reference genomes sharded, every GPU maps
                                                                         every query, hit rows gathered over NCCL in the library
    python bench_many.py --verify ROWS.json [--verify-shard R]           CPU only: regenerate the sampled queries and one shard
                                                                         of the references with numpy, map them with the CPU
                                                                         reference and compare the saved hit rows bit for bit

Prints one JSON line (and writes it to --out): genome-pairs/s and fragments/s with the queries resident in HBM, the
per-stage split with roofline fractions (SURVEY.md 8(d) algorithmic bytes over the CUDA-event stage timers), the realised
pairwise-identity histogram of the reported pairs, and -- single GPU -- the public API from host memory and the CPU
reference on a sample of the queries.  Every base is a pure function of (seed, tree node, position), so the collection of a
run can be regenerated without a GPU; nothing is read from /root/reference."""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench_workloads as W  # noqa: E402

FRAG = 3000
COUNTERS = ("fragments", "sketch_sum", "seeds", "candidates", "scanned", "events", "events_replayed", "mappings",
            "l1_sorted_fragments", "l1_small_fragments", "l2_fallback", "kernel_launches", "h2d_bytes", "d2h_bytes")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=256)
    ap.add_argument("--tree", default="genus", choices=["genus", "related"])
    ap.add_argument("--species", type=int, default=8, help="genus tree: strains per species")
    ap.add_argument("--genus", type=int, default=4, help="genus tree: species per genus")
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--scale", type=int, default=1, help="divide the genome lengths by this (quick tests)")
    ap.add_argument("--drafts", action="store_true", help="cut every genome into 200-500 contigs, half of them reverse-complemented (configs[2])")
    ap.add_argument("--cpu-sample", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--chunk", type=int, default=64, help="queries per library call (query_many)")
    ap.add_argument("--shard", default="queries", choices=["queries", "references"],
                    help="under torchrun: shard the queries (replicated index, configs[3]) or the reference genomes "
                         "(every rank maps all queries against its shard, hit rows all-gathered and merged, configs[4])")
    ap.add_argument("--sample-rows", type=int, default=16, help="queries (evenly spaced) whose hit rows are saved for --verify")
    ap.add_argument("--out", default=None, help="also write the JSON line (with the sampled rows) to this file")
    ap.add_argument("--verify", default=None, help="CPU only: check the sampled rows of this --out file against the CPU reference")
    ap.add_argument("--verify-shard", type=int, default=0, help="--verify: which reference shard to index on the CPU (-1: all genomes)")
    ap.add_argument("--verify-kind", default=None, choices=[None, "reference", "port"])
    ap.add_argument("--no-host", action="store_true", help="one GPU: skip the host-memory (e2e) legs")
    ap.add_argument("--profile", type=int, default=0, help="map this many queries between cudaProfilerStart/Stop and exit (ncu --profile-from-start off)")
    return ap.parse_args()


def collection(a):
    return W.Collection(a.tree, a.genomes, a.seed, species=a.species, genus=a.genus, scale=a.scale)


def workload_text(a, G):
    tree = ("genus / species / strain tree with independent genus roots (pairs of different genera are unrelated)" if a.tree == "genus"
            else "one root: family / genus / species / strain, every pair at 75-100 % identity")
    return "%d x %d synthetic genomes of 3-6 Mbp (%s), %s" % (
        G, G, "drafts of 200-500 contigs, strands flipped" if a.drafts else "complete genomes", tree)


def identity_histogram(rows_per_query, G):
    """Realised identity of the reported (query, reference) pairs, 1 % bins from 75 to 100, and the pairs not reported."""
    edges = np.arange(75, 101)
    hist = np.zeros(len(edges), dtype=np.int64)
    n = 0
    for q, rows in enumerate(rows_per_query):
        ident = np.asarray(rows["identity"], dtype=np.float64)
        n += len(ident)
        hist += np.bincount(np.clip(np.floor(ident).astype(np.int64) - 75, 0, len(edges) - 1), minlength=len(edges))
    return {"bins_percent_from": edges.tolist(), "pairs": hist.tolist(), "reported": int(n),
            "not_reported": int(len(rows_per_query) * G - n)}


def rows_of(hits, names_to_id=None):
    from pyfastani_b200 import sharding
    return sharding.hits_to_rows(hits, names_to_id) if names_to_id is not None else hits


def sample_ids(G, n):
    return sorted({int(round(x)) for x in np.linspace(0, G - 1, max(1, min(n, G)))})


def stage_report(inf, bases, peak):
    """Per-stage ms, algorithmic bytes and fraction of the HBM roofline from summed last_query_info counters."""
    from bench import stage_bytes
    alg = stage_bytes(inf, bases)
    out = {}
    for k, b in alg.items():
        ms = inf.get(k, 0.0)
        out[k] = {"ms": ms, "alg_bytes": b, "frac": (b / (ms * 1e-3) / 1e9 / peak) if ms > 0 else None}
    tot_ms = inf.get("ms_batch", 0.0) or inf.get("ms_total", 0.0)
    out["whole"] = {"ms": tot_ms, "alg_bytes": sum(alg.values()), "frac": sum(alg.values()) / max(tot_ms * 1e-3, 1e-12) / 1e9 / peak}
    return out


def add_info(acc, inf):
    for k, v in inf.items():
        if k.startswith("ms_") or k in COUNTERS:
            acc[k] = acc.get(k, 0) + v


# ------------------------------------------------------------------------------------------------------------------
# --verify: CPU only
# ------------------------------------------------------------------------------------------------------------------
def verify(a):
    from oracle.oracle import Oracle, available
    rec = json.load(open(a.verify))
    cfg = rec["config"]
    a.genomes, a.tree, a.seed, a.drafts = cfg["genomes"], cfg["tree"], cfg["seed"], cfg["drafts"]
    a.species, a.genus, a.scale = cfg.get("species", 8), cfg.get("genus", 4), cfg.get("scale", 1)
    col = collection(a)
    be = W.NumpyBackend()
    G = col.G
    offsets = rec.get("genome_offsets") or [0, G]
    lo, hi = (0, G) if a.verify_shard < 0 else (offsets[a.verify_shard], offsets[a.verify_shard + 1])
    kind = a.verify_kind or ("reference" if "reference" in available() else "port")
    orc = Oracle(kind)
    threads = os.cpu_count() or 1

    def genome(i):
        codes = col.codes(be, i)
        return W.contigs_numpy(codes, W.draft_plan(a.seed, i, len(codes)) if a.drafts else None)

    t0 = time.perf_counter()
    sk = orc.sketch()
    block = 64
    for b in range(lo, hi, block):
        ids = list(range(b, min(b + block, hi)))
        gs = [genome(i) for i in ids]                 # (in order: the ancestors of the tree are cached)
        if a.drafts or kind != "reference":
            for i, g in zip(ids, gs):
                sk.add_draft(i - lo, g)
        else:
            sk.add_genomes([i - lo for i in ids], [g[0] for g in gs], threads=threads)
    t_sketch = time.perf_counter() - t0
    t0 = time.perf_counter()
    sk.index()
    t_index = time.perf_counter() - t0
    checked = same = pairs = 0
    t_q = 0.0
    bad = []
    kw = {"threads": threads} if kind == "reference" else {}
    for q, rows in zip(rec["sample_queries"], rec["sample_rows"]):
        g = genome(q)
        t0 = time.perf_counter()
        oh, _ = sk.query_draft(g, **kw)
        t_q += time.perf_counter() - t0
        want = [(int(h["ref_genome"]) + lo, int(h["matches"]), int(h["fragments"]), float(np.float32(h["identity"]))) for h in oh]
        got = [(r[0], r[1], r[2], float(np.float32(r[3]))) for r in rows if lo <= r[0] < hi]
        checked += 1
        pairs += hi - lo
        if want == got:          # same rows in the same order (identity descending, ascending genome id among equals)
            same += 1
        else:
            bad.append({"query": q, "cpu": want[:5], "gpu": got[:5], "n_cpu": len(want), "n_gpu": len(got)})
    out = {"verify": a.verify, "kind": kind, "cores": threads if kind == "reference" else 1,
           "reference_genomes": [lo, hi], "queries_checked": checked, "queries_identical": same, "pairs_checked": pairs,
           "what": "hit rows (genome, matches, fragments, identity bits, order) of the sampled queries against genomes [%d, %d), "
                   "GPU run vs the CPU %s on inputs regenerated with numpy" % (lo, hi, kind),
           "cpu_seconds": {"sketch": t_sketch, "index": t_index, "queries": t_q},
           "cpu_pairs_per_s": pairs / max(t_q, 1e-9), "mismatches": bad[:4]}
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"))
    if same != checked:
        sys.exit(1)


# ------------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.verify:
        return verify(a)
    import torch
    import pyfastani_b200 as pf
    from bench import peaks
    from pyfastani_b200 import sharding

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = peaks()
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.tensor(list(b"TGCA"), dtype=torch.uint8, device=dev)

    def reduce(x, op="max", dtype=torch.float64):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=dtype, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.item()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- the collection, resident in HBM as ASCII contigs ------------------------------------------------------------
    col = collection(a)
    G = col.G
    be = W.TorchBackend(torch, dev)
    t0 = time.perf_counter()
    dev_contigs = []
    for i in range(G):
        codes = col.codes(be, i)
        dev_contigs.append(W.contigs_torch(torch, codes, W.draft_plan(a.seed, i, int(codes.shape[0])) if a.drafts else None, lut, comp))
    col._cache.clear()
    torch.cuda.synchronize(dev)
    t_generate = time.perf_counter() - t0
    lengths = [sum(int(p.numel()) for p in cs) for cs in dev_contigs]
    total_bp = sum(lengths)
    frag_counts = [sum(int(p.numel()) // FRAG for p in cs) for cs in dev_contigs]

    def wrap(cs):
        return [pf.DeviceSequence.from_pointer(p.data_ptr(), p.numel(), local, p) for p in cs]

    def item(cs):
        return cs if a.drafts else cs[0]

    # ---- index -------------------------------------------------------------------------------------------------
    by_refs = world > 1 and a.shard == "references"
    offsets = sharding.reference_shards(lengths, world) if by_refs else [0, G]
    my_refs = range(offsets[rank], offsets[rank + 1]) if by_refs else range(G)
    comm = None
    if by_refs:
        def exchange(uid):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if uid is not None:
                t.copy_(torch.from_numpy(np.frombuffer(uid, dtype=np.uint8).copy()))
            dist.broadcast(t, src=0)
            return t.cpu().numpy().tobytes()
        comm = sharding.connect(world, rank, local, exchange=exchange)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    sketch = pf.Sketch(device=local)
    names = [i - my_refs[0] for i in my_refs]          # local genome ids: what the row interface reports
    for b in range(0, len(names), 64):
        sketch.add_many(names[b:b + 64], [item(wrap(dev_contigs[i])) for i in my_refs[b:b + 64]])
    t_sketch = time.perf_counter() - t0
    sk_stats = sketch.build_stats
    n_min = len(sketch.minimizers)
    t0 = time.perf_counter()
    mapper = sketch.index()
    t_index = time.perf_counter() - t0
    ix_stats = mapper.build_stats
    build = {"generate_s": t_generate, "wall_sketch_s": t_sketch, "wall_index_s": t_index, "minimizers": n_min,
             "sketch_ms": sk_stats["ms_sketch"], "sketch_frac": 1.96 * sk_stats["bases"] / max(sk_stats["ms_sketch"] * 1e-3, 1e-12) / 1e9 / peak,
             "index_ms": ix_stats["ms_build"], "index_sort_ms": ix_stats["ms_sort"],
             "index_frac": 28.0 * n_min / max(ix_stats["ms_build"] * 1e-3, 1e-12) / 1e9 / peak}
    cfg = {"workload": None, "genomes": G, "tree": a.tree, "seed": a.seed, "drafts": bool(a.drafts), "species": a.species, "genus": a.genus, "scale": a.scale,
           "total_mbp": total_bp / 1e6, "chunk": a.chunk, "l2_policy": "inputs larger than L2 (index %.1f GB per GPU)" % (n_min * 50 / 1e9)}

    def finish(line, rows_all=None, rows_sample=None):
        line["config"] = cfg
        line["index_build"] = build
        line["peak"] = {"hbm_gbs": peak, "source": peak_src}
        if rows_all is not None:
            line["identity_histogram"] = identity_histogram(rows_all, G)
        if rows_sample is not None:
            line["sample_queries"] = [q for q, _ in rows_sample]
            line["sample_rows"] = [[[int(r["ref_genome"]), int(r["matches"]), int(r["fragments"]), float(r["identity"])] for r in rows]
                                   for _, rows in rows_sample]
        print(json.dumps({k: v for k, v in line.items() if k != "sample_rows"}))
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            json.dump(line, open(a.out, "w"))

    if world > 1:
        # ---- both multi-GPU layouts: every rank maps its list of queries in chunks through query_many -----------------
        if by_refs:
            # every rank maps every query against its shard (collective calls).  Neighbouring genomes of the collection are
            # related, so a chunk of consecutive queries would put all of its real mapping work on the one rank that holds
            # their relatives while the others wait at the chunk's all-gather: the queries go in a fixed shuffled order
            # (the same on every rank), which spreads every chunk over all shards.
            mine = [int(x) for x in np.random.default_rng(a.seed + 977).permutation(G)]
        else:
            mine = sharding.partition_queries(frag_counts, world)[rank]
        items = [item(wrap(dev_contigs[i])) for i in mine]
        n_calls = (G + a.chunk - 1) // a.chunk if by_refs else 0

        def run_all():
            acc, rows = {}, []
            for b in range(0, len(items), a.chunk):
                if by_refs:
                    rows += mapper.query_many(items[b:b + a.chunk], rows=True, comm=comm, genome_offsets=offsets)
                else:
                    rows += mapper.query_many(items[b:b + a.chunk], rows=True)
                add_info(acc, mapper.last_query_info)
            return acc, rows

        if by_refs:
            mapper.query_many(items[:min(8, len(items))], rows=True, comm=comm, genome_offsets=offsets)
        else:
            mapper.query_many(items[:min(8, len(items))], rows=True)
        best = None
        for _ in range(a.repeat):
            barrier()
            t0 = time.perf_counter()
            acc, rows = run_all()
            barrier()
            t = reduce(time.perf_counter() - t0)
            if best is None or t < best[0]:
                best = (t, acc, rows)
        t_all, acc, rows = best
        # per-rank device time and counters -> max / sum over the ranks
        stage_max = {k: reduce(float(v)) for k, v in sorted(acc.items()) if k.startswith("ms_")}
        stage_min = {k: -reduce(-float(v)) for k, v in sorted(acc.items()) if k.startswith("ms_")}
        count_sum = {k: int(reduce(int(acc.get(k, 0)), "sum", torch.int64)) for k in COUNTERS}
        if by_refs:
            rows_all = [None] * G                     # merged rows of every query, identical on every rank
            for q, r in zip(mine, rows):
                rows_all[q] = r
            ok = all(len(m) and q in m["ref_genome"][:4] for q, m in enumerate(rows_all))    # every genome finds itself at the top
        else:
            # gather the rows of the partitioned queries on rank 0 for the histogram / sample (outside the timed region)
            mine_rows = [(q, r) for q, r in zip(mine, rows)]
            gathered = [None] * world
            dist.all_gather_object(gathered, [(q, r.tolist()) for q, r in mine_rows])
            rows_all = [None] * G
            for part in gathered:
                for q, r in part:
                    rows_all[q] = np.array([tuple(x) for x in r], dtype=sharding.HIT_DT)
            ok = all(len(m) and q in m["ref_genome"][:4] for q, m in enumerate(rows_all))
        assert ok, "a genome does not find itself among its best hits"
        if rank == 0:
            inf = dict(count_sum)
            inf.update(stage_max)
            bases = total_bp * (world if by_refs else 1)          # reference-sharded: every rank sketches every query
            stages = stage_report(inf, bases, peak * world)
            cfg["workload"] = ("configs[%d] layout: %s; %s" % (
                4 if by_refs else 3, workload_text(a, G),
                ("reference genomes sharded over %d GPUs (%s per rank), every rank maps all queries (resident, in one shuffled order), per %d queries one library call "
                 "(fa_query_batch_sharded): mapping + one ncclAllGather of the hit rows + merge" % (world, [offsets[r + 1] - offsets[r] for r in range(world)], a.chunk))
                if by_refs else "queries dealt to %d GPUs (LPT by fragments), replicated index, query_many per %d queries" % (world, a.chunk)))
            line = {"metric": "genome_pairs_per_s", "unit": "genome-pairs/s", "n_gpus": world, "scaling": "strong",
                    "value": G * G / t_all, "fragments_per_s": sum(frag_counts) / t_all, "ms_per_query": t_all / G * 1e3,
                    "wall_s": t_all, "timing": "wall clock between barriers (+ cuda synchronize) around all %d queries, max over ranks, best of %d" % (G, a.repeat),
                    "hits": int(sum(len(m) for m in rows_all)),
                    "stages_max_over_ranks": stages, "stages_ms_min_over_ranks": stage_min,
                    "device_busy_frac": stage_max.get("ms_batch", 0.0) * 1e-3 / t_all,
                    "counters_sum_over_ranks": count_sum, "genome_offsets": offsets,
                    "gather": dict(comm.info, calls=n_calls) if comm is not None else None}
            finish(line, rows_all, [(q, rows_all[q]) for q in sample_ids(G, a.sample_rows)])
        dist.destroy_process_group()
        return

    # ---- one GPU: resident queries, one call per query (library timers) ----------------------------------------------
    cfg["workload"] = "configs[%d] shape: %s" % (2 if a.drafts else 3, workload_text(a, G))
    dq = [item(wrap(cs)) for cs in dev_contigs]

    def one(q):
        return mapper.query_draft(q) if a.drafts else mapper.query_genome(q)

    for q in dq[:4]:
        one(q)
    if a.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        if a.profile == 1:
            one(dq[4])
        else:               # one query_many call: a pass of one query, then shared passes
            mapper.query_many(dq[4:4 + a.profile])
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return
    ident = {n: n for n in names}
    best_one = None
    for _ in range(a.repeat):
        torch.cuda.synchronize(dev)
        tw = time.perf_counter()
        acc, res = {}, []
        for q in dq:
            res.append(one(q))
            add_info(acc, mapper.last_query_info)
        wall = time.perf_counter() - tw
        if best_one is None or acc["ms_total"] < best_one[0]["ms_total"]:
            best_one = (acc, wall, res)
    acc_one, wall_one, res_one = best_one
    rows_one = [sharding.hits_to_rows(h, ident) for h in res_one]

    # ---- resident queries, query_many in chunks: light queries share passes of the pipeline --------------------------
    best_many = None
    for _ in range(a.repeat):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        acc, rows = {}, []
        for b in range(0, G, a.chunk):
            rows += mapper.query_many(dq[b:b + a.chunk], rows=True)
            add_info(acc, mapper.last_query_info)
        dt = time.perf_counter() - t0
        if best_many is None or acc["ms_batch"] < best_many[0]["ms_batch"]:
            best_many = (acc, dt, rows)
    acc_many, wall_many, rows_many = best_many
    assert all(np.array_equal(x, y) for x, y in zip(rows_one, rows_many)), "query_many rows differ from one call per query"

    # ---- from host memory: the public API, wall clock ------------------------------------------------------------
    e2e = e2e_many = None
    host = None
    if not a.no_host or a.cpu_sample > 0:
        host = [[p.cpu().numpy().tobytes() for p in cs] for cs in dev_contigs]
    if not a.no_host:
        hq = [item(cs) for cs in host]
        for q in hq[:4]:
            one(q)
        for _ in range(a.repeat):
            t0 = time.perf_counter()
            res_host = [one(q) for q in hq]
            dt = time.perf_counter() - t0
            e2e = dt if e2e is None else min(e2e, dt)
        for _ in range(a.repeat):
            t0 = time.perf_counter()
            rows_host = []
            for b in range(0, G, a.chunk):
                rows_host += mapper.query_many(hq[b:b + a.chunk], rows=True)
            dt = time.perf_counter() - t0
            e2e_many = dt if e2e_many is None else min(e2e_many, dt)
        assert all(np.array_equal(sharding.hits_to_rows(x, ident), y) and np.array_equal(y, z) for x, y, z in zip(res_host, rows_one, rows_host))

    # ---- CPU reference on a sample of the queries (whole index) --------------------------------------------------
    cpu = None
    if a.cpu_sample > 0:
        from oracle.oracle import Oracle, available
        kind = "reference" if "reference" in available() else "port"
        orc = Oracle(kind)
        threads = (os.cpu_count() or 1) if kind == "reference" else 1
        sk = orc.sketch()
        t0 = time.perf_counter()
        if a.drafts or kind != "reference":
            for i, cs in enumerate(host):
                sk.add_draft(i, cs)
        else:
            for b in range(0, G, 64):
                sk.add_genomes(list(range(b, min(b + 64, G))), [cs[0] for cs in host[b:b + 64]], threads=threads)
        t_cpu_sketch = time.perf_counter() - t0
        t0 = time.perf_counter()
        sk.index()
        t_cpu_index = time.perf_counter() - t0
        kw = {"threads": threads} if kind == "reference" else {}
        ids = sample_ids(G, a.cpu_sample)
        t_cpu, same = 0.0, 0
        for i in ids:
            t0 = time.perf_counter()
            oh, _ = sk.query_draft(host[i], **kw)
            t_cpu += time.perf_counter() - t0
            same += int(np.array_equal(oh, rows_one[i]))
        pool = None
        if kind == "reference" and len(ids) > 1:        # BASELINE.md section 3, schedule (ii): a thread pool over queries, one thread each
            qs = [ids[i % len(ids)] for i in range(min(threads, 4 * len(ids)))]
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=min(threads, len(qs))) as ex:
                list(ex.map(lambda i: sk.query_draft(host[i], threads=1), qs))
            pool = len(qs) * G / (time.perf_counter() - t0)
        cpu = {"value": max(len(ids) * G / t_cpu, pool or 0.0), "unit": "genome-pairs/s", "cores": threads, "kind": kind,
               "sample": "%d of the %d queries against the whole index, %.2f s per query with the intra-query pool (sketch %.1f s, index %.1f s)"
                         % (len(ids), G, t_cpu / len(ids), t_cpu_sketch, t_cpu_index),
               "schedules": {"intra_query_pool_T%d" % threads: len(ids) * G / t_cpu, "thread_pool_over_queries_T1_each": pool},
               "queries_identical": same, "queries_checked": len(ids), "pairs_checked": len(ids) * G}
        assert same == len(ids), "GPU hit rows differ from the CPU reference"

    pairs = G * G
    frags = sum(frag_counts)
    line = {
        "metric": "genome_pairs_per_s", "unit": "genome-pairs/s", "n_gpus": 1,
        "value": pairs / (acc_many["ms_batch"] * 1e-3), "ms_per_query": acc_many["ms_batch"] / G,
        "fragments_per_s": frags / (acc_many["ms_batch"] * 1e-3),
        "timing": "query_many in chunks of %d, queries resident in HBM, one CUDA-event pair per call on the library's stream" % a.chunk,
        "wall": {"value": pairs / wall_many, "ms_per_query": wall_many / G * 1e3},
        "stages": stage_report(acc_many, total_bp, peak),
        "gpu_launches_per_query": acc_many["kernel_launches"] / G,
        "one_call_per_query": {"value": pairs / (acc_one["ms_total"] * 1e-3), "ms_per_query": acc_one["ms_total"] / G,
                               "wall_value": pairs / wall_one, "stages_ms_per_query": {k: v / G for k, v in sorted(acc_one.items()) if k.startswith("ms_")},
                               "gpu_launches_per_query": acc_one["kernel_launches"] / G},
        "e2e": None if e2e is None else {"value": pairs / e2e_many, "ms_per_query": e2e_many / G * 1e3, "h2d_bytes_per_query": total_bp // G,
                                         "one_call_per_query": {"value": pairs / e2e, "ms_per_query": e2e / G * 1e3}},
        "counters": {k: acc_many.get(k, 0) for k in COUNTERS}, "hits": int(sum(len(r) for r in rows_many)),
        "cpu_baseline": cpu, "genome_offsets": [0, G],
    }
    finish(line, rows_many, [(q, rows_many[q]) for q in sample_ids(G, a.sample_rows)])


if __name__ == "__main__":
    main()
