#!/usr/bin/env python
"""Secondary measurement (not the driver's contract line, that is bench.py): the many-to-many shape of BASELINE
configs[3] at a size one GPU sets up in seconds -- G synthetic genomes of 3-6 Mbp in a genus / species / strain tree
(pairwise identity 75-100 %), every genome mapped against the index of all of them.

    python bench_many.py [--genomes 256] [--drafts] [--cpu-sample 2]

Prints one JSON line: genome-pairs/s with the queries resident in HBM (the library's own CUDA-event timers), through
the public API from host memory (wall clock, one `query_genome` / `query_draft` per genome and one `query_many` over
all of them), the per-stage split, and -- for `--cpu-sample` queries -- the CPU reference on the same inputs with a
bit-exact comparison of the hit rows.  Genomes are generated on the GPU; nothing is read from /root/reference."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
FRAG = 3000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=256)
    ap.add_argument("--species", type=int, default=8, help="strains per species")
    ap.add_argument("--genus", type=int, default=4, help="species per genus")
    ap.add_argument("--seed", type=int, default=4)
    ap.add_argument("--drafts", action="store_true", help="cut every genome into 200-500 contigs, half of them reverse-complemented (configs[2])")
    ap.add_argument("--cpu-sample", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--shard", default="queries", choices=["queries", "references"],
                    help="under torchrun: shard the queries (replicated index, configs[3]) or the reference genomes "
                         "(every rank maps all queries against its shard, hit rows all-gathered and merged, configs[4])")
    ap.add_argument("--profile", type=int, default=0, help="map this many queries between cudaProfilerStart/Stop and exit (ncu --profile-from-start off)")
    return ap.parse_args()


def mutate(torch, codes, ident, g):
    hit = torch.rand(codes.shape, generator=g, device=codes.device) < (1.0 - ident)
    shift = torch.randint(1, 4, codes.shape, generator=g, device=codes.device, dtype=torch.uint8)
    return (codes + hit.to(torch.uint8) * shift) & 3


def main():
    a = parse()
    import torch
    import pyfastani_b200 as pf

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(a.seed)
    g = torch.Generator(device=dev)
    g.manual_seed(a.seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.tensor(list(b"TGCA"), dtype=torch.uint8, device=dev)

    # ---- the tree: genus root -> species (0.80-0.90 of the root) -> strains (0.95-0.999 of the species) ----------
    genomes = []            # uint8 code tensors on the device
    per_genus = a.genus * a.species
    while len(genomes) < a.genomes:
        length = int(rng.integers(3_000_000, 6_000_001))
        root = torch.randint(0, 4, (length,), generator=g, device=dev, dtype=torch.uint8)
        for _ in range(a.genus):
            sp = mutate(torch, root, float(rng.uniform(0.80, 0.90)), g)
            for _ in range(a.species):
                if len(genomes) < a.genomes:
                    genomes.append(mutate(torch, sp, float(rng.uniform(0.95, 0.999)), g))
        del root
    G = len(genomes)

    def as_contigs(codes):
        """ASCII contigs of one genome on the device (one contig, or a fragmented, strand-flipped, permuted draft)."""
        seq = lut[codes.long()]
        if not a.drafts:
            return [seq]
        n = int(rng.integers(200, 501))
        cuts = np.sort(rng.choice(np.arange(1000, codes.numel() - 1000), size=n - 1, replace=False))
        bounds = np.concatenate([[0], cuts, [codes.numel()]])
        parts = []
        for i in rng.permutation(n):
            p = seq[int(bounds[i]):int(bounds[i + 1])]
            if rng.random() < 0.5:
                p = comp[codes[int(bounds[i]):int(bounds[i + 1])].long()].flip(0)
            parts.append(p.contiguous())
        return parts

    dev_contigs = [as_contigs(c) for c in genomes]
    del genomes
    torch.cuda.synchronize(dev)
    total_bp = sum(int(p.numel()) for cs in dev_contigs for p in cs)

    def wrap(cs):
        return [pf.DeviceSequence.from_pointer(p.data_ptr(), p.numel(), local, p) for p in cs]

    # ---- index -------------------------------------------------------------------------------------------------
    by_refs = world > 1 and a.shard == "references"
    if by_refs:
        from pyfastani_b200 import sharding
        offsets = sharding.reference_shards([sum(int(p.numel()) for p in cs) for cs in dev_contigs], world)
        my_refs = range(offsets[rank], offsets[rank + 1])
    else:
        my_refs = range(len(dev_contigs))
    t0 = time.perf_counter()
    sketch = pf.Sketch(device=local)
    for i in my_refs:                       # (names are the local genome ids: what query_reference_sharded expects)
        if a.drafts:
            sketch.add_draft(i - my_refs[0], wrap(dev_contigs[i]))
        else:
            sketch.add_genome(i - my_refs[0], wrap(dev_contigs[i])[0])
    t_sketch = time.perf_counter() - t0
    n_min = len(sketch.minimizers)
    t0 = time.perf_counter()
    mapper = sketch.index()
    t_index = time.perf_counter() - t0

    def one(q):
        return mapper.query_draft(q) if a.drafts else mapper.query_genome(q[0])

    if by_refs:
        # ---- configs[4] shape: reference genomes sharded, every rank maps ALL queries against its shard (query_many),
        # the hit rows are all-gathered over NCCL and merged into the global order on every rank -----------------------
        items = [cs if a.drafts else cs[0] for cs in (wrap(c) for c in dev_contigs)]
        chunk = 64

        def run_all():
            merged = []
            for b in range(0, len(items), chunk):
                merged += sharding.query_reference_sharded(mapper, items[b:b + chunk], offsets, device=dev, drafts=a.drafts)
            return merged

        run_all()
        best = None
        for _ in range(a.repeat):
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            merged = run_all()
            torch.cuda.synchronize(dev)
            dist.barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        assert all(len(m) and q in m["ref_genome"][:4] for q, m in enumerate(merged))     # every genome finds itself at the top
        if rank == 0:
            G = len(dev_contigs)
            print(json.dumps({
                "metric": "genome_pairs_per_s", "unit": "genome-pairs/s", "n_gpus": world, "scaling": "strong",
                "config": {"workload": "configs[4] shape, scaled: %d x %d synthetic genomes of 3-6 Mbp, reference genomes sharded over %d GPUs "
                                       "(%s per rank), every rank maps all queries (resident), hit rows all-gathered per %d queries"
                                       % (G, G, world, [offsets[r + 1] - offsets[r] for r in range(world)], chunk),
                           "genomes": G, "index_minimizers_rank0": n_min, "seed": a.seed},
                "value": G * G / best, "ms_per_query": best / G * 1e3, "hits": int(sum(len(m) for m in merged)),
                "index_build": {"sketch_s": t_sketch, "index_s": t_index}}))
        dist.destroy_process_group()
        return

    if world > 1:
        # ---- configs[3] proper: queries sharded over the GPUs (LPT by fragment count), index replicated, no collective
        # in the mapping path; every rank maps its share through query_many, times are the max over ranks ----------------
        from pyfastani_b200 import sharding
        frag_counts = [sum(int(p.numel()) // FRAG for p in cs) for cs in dev_contigs]
        mine = sharding.partition_queries(frag_counts, world)[rank]
        dq = [wrap(dev_contigs[i]) for i in mine]
        host = [[p.cpu().numpy().tobytes() for p in dev_contigs[i]] for i in mine]

        def timed(items):
            mapper.query_many(items[:3])
            best = None
            for _ in range(a.repeat):
                torch.cuda.synchronize(dev)
                dist.barrier()
                t0 = time.perf_counter()
                res = mapper.query_many(items)
                torch.cuda.synchronize(dev)
                dist.barrier()
                t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = float(t.item()) if best is None else min(best, float(t.item()))
            return best, res

        t_res, r1 = timed([q if a.drafts else q[0] for q in dq])
        t_host, r2 = timed([q if a.drafts else q[0] for q in host])
        assert [[(h.name, h.matches, h.identity) for h in x] for x in r1] == [[(h.name, h.matches, h.identity) for h in x] for x in r2]
        n_hits = torch.tensor([sum(len(x) for x in r1)], dtype=torch.int64, device=dev)
        dist.all_reduce(n_hits)
        if rank == 0:
            pairs = len(dev_contigs) ** 2
            print(json.dumps({
                "metric": "genome_pairs_per_s", "unit": "genome-pairs/s", "n_gpus": world, "scaling": "strong",
                "config": {"workload": "configs[%d] shape, scaled: %d x %d synthetic genomes of 3-6 Mbp, queries sharded over %d GPUs "
                                       "(LPT by fragments), replicated index, query_many per rank" % (2 if a.drafts else 3, len(dev_contigs), len(dev_contigs), world),
                           "genomes": len(dev_contigs), "index_minimizers": n_min, "seed": a.seed},
                "value": pairs / t_res, "ms_per_query": t_res / len(dev_contigs) * 1e3,
                "e2e": {"value": pairs / t_host, "ms_per_query": t_host / len(dev_contigs) * 1e3, "h2d_bytes_per_query": total_bp // len(dev_contigs)},
                "hits": int(n_hits.item()), "queries_per_rank": len(mine), "index_build": {"sketch_s": t_sketch, "index_s": t_index}}))
        dist.destroy_process_group()
        return

    # ---- resident queries: library timers ------------------------------------------------------------------------
    dq = [wrap(cs) for cs in dev_contigs]
    for q in dq[:4]:
        one(q)
    if a.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        if a.profile == 1:
            one(dq[4])
        else:               # one query_many call: a pass of one query, then shared passes
            mapper.query_many([q if a.drafts else q[0] for q in dq[4:4 + a.profile]])
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return
    stage, best_dev, best_wall = {}, None, None
    for _ in range(a.repeat):
        torch.cuda.synchronize(dev)
        tw = time.perf_counter()
        ms, st, counters, launches = 0.0, {}, {}, 0
        res_dev = []
        for q in dq:
            res_dev.append(one(q))
            inf = mapper.last_query_info
            ms += inf["ms_total"]
            launches += inf["kernel_launches"]
            for k, v in inf.items():
                if k.startswith("ms_"):
                    st[k] = st.get(k, 0.0) + v
                elif k in ("fragments", "seeds", "candidates", "scanned", "events", "events_replayed", "mappings", "l1_sorted_fragments"):
                    counters[k] = counters.get(k, 0) + v
        wall = time.perf_counter() - tw
        if best_dev is None or ms < best_dev:
            best_dev, stage, best_wall = ms, st, wall

    # ---- resident queries, one query_many call: light queries share passes of the pipeline -------------------------
    many_dev = None
    for _ in range(a.repeat):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res_many_dev = mapper.query_many([q if a.drafts else q[0] for q in dq])
        dt = time.perf_counter() - t0
        inf = mapper.last_query_info
        if many_dev is None or inf["ms_total"] < many_dev[0]:
            many_dev = (inf["ms_total"], dt, inf["kernel_launches"], {k: v / len(dq) for k, v in sorted(inf.items()) if k.startswith("ms_")})

    # ---- from host memory: the public API, wall clock ------------------------------------------------------------
    host = [[p.cpu().numpy().tobytes() for p in cs] for cs in dev_contigs]
    for q in host[:4]:
        one(q)
    e2e = None
    for _ in range(a.repeat):
        t0 = time.perf_counter()
        res_host = [one(q) for q in host]
        dt = time.perf_counter() - t0
        e2e = dt if e2e is None else min(e2e, dt)
    many = None
    for _ in range(a.repeat):
        t0 = time.perf_counter()
        res_many = mapper.query_many([q if a.drafts else q[0] for q in host])
        dt = time.perf_counter() - t0
        many = dt if many is None else min(many, dt)

    def rows(hs):
        return [(h.name, h.matches, h.fragments, h.identity) for h in hs]

    assert all(rows(x) == rows(y) == rows(z) == rows(u) for x, y, z, u in zip(res_dev, res_host, res_many, res_many_dev))
    n_hits = sum(len(h) for h in res_dev)

    # ---- CPU reference on a sample of the queries (whole index) --------------------------------------------------
    cpu = None
    if a.cpu_sample > 0:
        from oracle.oracle import Oracle, available
        kind = "reference" if "reference" in available() else "port"
        orc = Oracle(kind)
        sk = orc.sketch()
        for i, cs in enumerate(host):
            sk.add_draft(i, cs) if a.drafts else sk.add_genome(i, cs[0])
        t0 = time.perf_counter()
        sk.index()
        t_cpu_index = time.perf_counter() - t0
        threads = (os.cpu_count() or 1) if kind == "reference" else 1
        kw = {"threads": threads} if kind == "reference" else {}
        ids = sorted({int(round(x)) for x in np.linspace(0, G - 1, a.cpu_sample)})
        t_cpu, same = 0.0, 0
        for i in ids:
            t0 = time.perf_counter()
            oh, _ = (sk.query_draft(host[i], **kw) if a.drafts else sk.query_genome(host[i][0], **kw))
            t_cpu += time.perf_counter() - t0
            want = [(int(h["ref_genome"]), int(h["matches"]), int(h["fragments"]), float(np.float32(h["identity"]))) for h in oh]
            got = [(h.name, h.matches, h.fragments, float(np.float32(h.identity))) for h in res_dev[i]]
            same += int(sorted(want) == sorted(got))
        cpu = {"value": len(ids) * G / t_cpu, "unit": "genome-pairs/s", "cores": threads, "kind": kind,
               "sample": "%d of the %d queries against the whole index, %.2f s per query (index build %.1f s)"
                         % (len(ids), G, t_cpu / len(ids), t_cpu_index),
               "queries_identical": same, "queries_checked": len(ids)}
        assert same == len(ids), "GPU hit rows differ from the CPU reference"

    pairs = G * G
    frags = counters.get("fragments", 0)
    line = {
        "metric": "genome_pairs_per_s", "unit": "genome-pairs/s", "n_gpus": 1,
        "config": {"workload": "configs[%d] shape, scaled: %d x %d synthetic genomes of 3-6 Mbp (%s), genus/species/strain tree"
                               % (2 if a.drafts else 3, G, G, "drafts of 200-500 contigs, strands flipped" if a.drafts else "complete genomes"),
                   "genomes": G, "total_mbp": total_bp / 1e6, "index_minimizers": n_min, "seed": a.seed},
        "value": pairs / (best_dev * 1e-3), "ms_per_query": best_dev / G, "fragments_per_s": frags / (best_dev * 1e-3),
        "resident_wall": {"value": pairs / best_wall, "ms_per_query": best_wall / G * 1e3},
        "resident_query_many": {"value": pairs / (many_dev[0] * 1e-3), "ms_per_query": many_dev[0] / G,
                                "wall_value": pairs / many_dev[1], "wall_ms_per_query": many_dev[1] / G * 1e3,
                                "gpu_launches_per_query": many_dev[2] / G, "stages_ms_per_query": many_dev[3]},
        "e2e": {"value": pairs / e2e, "ms_per_query": e2e / G * 1e3, "h2d_bytes_per_query": total_bp // G},
        "e2e_query_many": {"value": pairs / many, "ms_per_query": many / G * 1e3},
        "stages_ms_per_query": {k: v / G for k, v in sorted(stage.items())},
        "counters": counters, "gpu_launches_per_query": launches / G, "hits": n_hits,
        "index_build": {"sketch_s": t_sketch, "index_s": t_index},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
