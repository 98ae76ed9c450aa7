#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (genome-pairs/s, fragments mapped/s, % of HBM roofline) on
BASELINE.json configs[1]: one synthetic 5 Mbp query genome against an index of 1,000 synthetic
5 Mbp references mutated to 80-99 % identity, on N B200s (replicated index, one query stream per
GPU: weak scaling, no collective in the mapping path).

A step = one pass of the hot path over one query genome:  Mapper.query_genome(query)  against the
resident index  =  n_refs genome pairs, 1,666 fragments.  The index build (sketch all references +
index) is set-up, as in the reference's own benchmark (benches/mapping/bench.py:34-66), and is
reported beside the metric.

  python bench.py [--gpus N --steps K --warmup W]           our arm (one JSON line)
  python bench.py --impl reference [...]                    the reference's CPU code on the host cores

`value` times K steps with the query already resident in HBM (CUDA events on the library's
stream, max over ranks); `e2e` times the same K steps through pyfastani_b200's public API with
HOST buffers (pinned staging, H2D and D2H inside the timed region, wall clock).
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
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC, UNIT = "genome_pairs_per_s", "genome-pairs/s"
FRAG = 3000


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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refs", type=int, default=1000)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--cpu-sample-refs", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="queries", choices=["queries", "references"],
                    help="N > 1: 'queries' = replicated index, one query stream per GPU (weak scaling, the default); "
                         "'references' = whole reference genomes split over the GPUs, every GPU maps the same query, hit rows "
                         "all-gathered over NCCL and merged (SURVEY.md 8(e), BASELINE config 5 layout; strong scaling)")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the device-timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_traffic(kernel, a):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` capture
    of this workload (profiles/traffic.json, written by profiles/summarize_ncu.py); None for any other workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t.get("refs") == a.refs and t.get("length") == a.length:
            return t["kernels"].get(kernel)
    except Exception:
        pass
    return None


def workload_name(a):
    return ("configs[1] 1-to-many: one synthetic %.1f Mbp query vs %d synthetic %.1f Mbp references at 80-99%% identity"
            % (a.length / 1e6, a.refs, a.length / 1e6))


# ---------------------------------------------------------------------------------------------
# synthetic data: the generator of tests/synth.py, with the per-reference
# mutation drawn on the GPU so that setting up 5 Gbp of references takes seconds, not minutes
# ---------------------------------------------------------------------------------------------
def base_codes(a):
    return np.random.default_rng(a.seed).integers(0, 4, size=a.length, dtype=np.uint8)


def identities(a):
    return np.linspace(0.80, 0.99, a.refs)


def reference_on_device(torch, base_dev, lut, ident, seed, index):
    g = torch.Generator(device=base_dev.device)
    g.manual_seed(seed * 1_000_003 + index)
    hit = torch.rand(base_dev.shape, generator=g, device=base_dev.device) < (1.0 - ident)
    shift = torch.randint(1, 4, base_dev.shape, generator=g, device=base_dev.device, dtype=torch.uint8)
    codes = (base_dev + hit.to(torch.uint8) * shift) & 3
    return lut[codes.long()]


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
# CPU baseline: the reference's own code (oracle/_ref) or, if absent, the C port
# ---------------------------------------------------------------------------------------------
def cpu_arm(a, query, sample_refs, steps, warmup):
    """Index `sample_refs` on the host, then time `steps` queries with all host threads."""
    from oracle.oracle import Oracle, available
    kind = "reference" if "reference" in available() else "port"
    orc = Oracle(kind)
    cores = os.cpu_count() or 1
    threads = cores if kind == "reference" else 1
    sk = orc.sketch()
    for i, r in sample_refs:
        sk.add_genome(i, r)
    sk.index()
    kw = {"threads": threads} if kind == "reference" else {}
    hits = None
    for _ in range(warmup):
        hits, _info = sk.query_genome(query, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        hits, _info = sk.query_genome(query, **kw)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    pairs = len(sample_refs)
    return {"value": pairs / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "the full %.1f Mbp query vs %d of the %d references (evenly spaced identities), index prebuilt, %.2f s per query"
                      % (a.length / 1e6, pairs, a.refs, dt),
            "fragments_per_s": (a.length // FRAG) / dt, "s_per_query": dt}, hits, [i for i, _ in sample_refs]


def sample_ids(a):
    n = max(1, min(a.cpu_sample_refs, a.refs))
    return sorted({int(round(x)) for x in np.linspace(0, a.refs - 1, n)})


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the same path on the host cores,
    each step a bounded sample of the workload (full query, a subset of the references)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    base = base_codes(a)
    query = synth.to_bytes(base)
    rng_refs = []
    idents = identities(a)
    for i in sample_ids(a):
        rng = np.random.default_rng(a.seed * 1_000_003 + i)
        rng_refs.append((i, synth.to_bytes(synth.mutate_codes(rng, base, float(idents[i])))))
    steps, warm = max(1, min(a.steps, 5)), max(0, min(a.warmup, 1))
    cb, _, _ = cpu_arm(a, query, rng_refs, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": cb["s_per_query"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "refs": a.refs, "length": a.length, "fragment_length": FRAG, "k": 16},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "fragments_per_s": cb["fragments_per_s"],
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    import pyfastani_b200 as pf
    import synth

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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- set-up: references generated in HBM, sketched and indexed (not in the timed region) ----
    base = base_codes(a)
    query = synth.to_bytes(base)
    base_dev = torch.from_numpy(base).to(dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    idents = identities(a)
    keep = set(sample_ids(a)) if (rank == 0 and world == 1 and not a.no_cpu_baseline) else set()
    sample_refs = []
    sketch = pf.Sketch(device=local)
    by_refs = a.shard == "references" and world > 1
    from pyfastani_b200 import sharding
    offsets = sharding.reference_shards([a.length] * a.refs, world) if by_refs else [0, a.refs]
    my_refs = range(offsets[rank], offsets[rank + 1]) if by_refs else range(a.refs)
    shard_cap = max(offsets[r + 1] - offsets[r] for r in range(len(offsets) - 1))      # most hit rows one rank can hold
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in my_refs:
        ref = reference_on_device(torch, base_dev, lut, float(idents[i]), a.seed, i)
        torch.cuda.synchronize(dev)
        sketch.add_genome(i, pf.DeviceSequence.from_pointer(ref.data_ptr(), ref.numel(), local, ref))
        if i in keep:
            sample_refs.append((i, ref.cpu().numpy().tobytes()))
        del ref
    t_sketch = time.perf_counter() - t0
    n_min = len(sketch.minimizers)
    t0 = time.perf_counter()
    mapper = sketch.index()
    t_index = time.perf_counter() - t0
    del base_dev
    torch.cuda.empty_cache()

    q_dev = pf.DeviceSequence.from_host(query, local)
    frags = a.length // FRAG
    pairs = a.refs

    # ---- warm-up ------------------------------------------------------------------------------
    hits = None
    name_to_local = {n: j for j, n in enumerate(mapper.names)}

    def map_query(q):
        """One step.  Reference-sharded: this rank's hits (local genome ids) are all-gathered as 16-byte rows over
        NCCL and merged into the global order on every rank."""
        hs = mapper.query_genome(q)
        if not by_refs:
            return hs
        rows = sharding.hits_to_rows(hs, name_to_local)
        return sharding.merge_hits(sharding.gather_hits([rows], device=dev, cap=shard_cap)[0], offsets)

    for _ in range(max(a.warmup, 1)):
        hits = map_query(q_dev)
        map_query(query)
    info0 = dict(mapper.last_query_info)

    # ---- timed: K steps, query resident in HBM (CUDA events inside the library) -----------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    stage = {}
    dev_ms = 0.0
    launches = 0
    if a.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
    t_wall = time.perf_counter()
    for _ in range(a.steps):
        map_query(q_dev)
        inf = mapper.last_query_info
        dev_ms += inf["ms_total"]
        launches += inf["kernel_launches"]
        for k, v in inf.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    if a.profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
    barrier()
    if by_refs:     # the gather is part of the step: wall clock between the barriers instead of the library's kernel timers
        dev_ms = (time.perf_counter() - t_wall) * 1e3
    dev_ms = max_over_ranks(dev_ms)
    inf = dict(mapper.last_query_info)

    # ---- timed: the same K steps through the public API with host buffers (wall clock) ----------
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(a.steps):
        hits_e2e = map_query(query)
        h2d += mapper.last_query_info["h2d_bytes"]
        d2h += mapper.last_query_info["d2h_bytes"]
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.summary()
    if by_refs:
        assert np.array_equal(hits_e2e, hits) and len(hits) <= a.refs
        assert np.all(np.diff(hits["identity"].astype(np.float64)) <= 0)
    else:
        assert [(h.name, h.matches, h.identity) for h in hits_e2e] == [(h.name, h.matches, h.identity) for h in hits]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (SURVEY.md 8(d) algorithmic bytes, realised counters) ---
    peak, which = peaks()
    s_mean = inf["sketch_sum"] / max(inf["fragments"], 1)
    alg = {
        "ms_sketch": 1.96 * a.length,
        "ms_lookup": 36.0 * inf["sketch_sum"],
        # L1 on chip (l1_fused_kernel): the position lists once (4 B / seed), one 4-byte gpos gather per seed, 16 B per
        # candidate; ms_seed_sort is the device-wide sort of the fragments that do not fit on chip (none here)
        "ms_seed_sort": 16.0 * inf["seeds"] * inf["l1_sorted_fragments"] / max(inf["fragments"], 1),
        "ms_l1": 8.0 * inf["seeds"] + 16.0 * inf["candidates"],
        # L2 = prep (index searches) + events (classify + merge: reads the (hash, wpos) stream once, writes
        # 2-byte events) + slide (replays the events up to its early stop, writes 16-byte results)
        "ms_l2_prep": 56.0 * inf["candidates"],
        # events: the (hash, order word) stream once (8 B / element), 2-byte events out
        "ms_l2_events": 8.0 * inf["scanned"] + 2.0 * inf["events"] + (4.0 * s_mean + 16.0) * inf["candidates"],
        "ms_l2_slide": 2.0 * inf["events_replayed"] + 34.0 * inf["candidates"],
        "ms_cgi": 16.0 * inf["candidates"],
    }
    per_step = {k: v / a.steps for k, v in stage.items()}
    top = max(alg, key=lambda k: per_step.get(k, 0.0))
    top_ms = per_step[top]
    achieved = alg[top] / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
    kernel_names = {"ms_sketch": "sketch_kernel", "ms_lookup": "lookup_kernel", "ms_seed_sort": "fill_seeds+DeviceRadixSort",
                    "ms_l1": "l1_fused_kernel", "ms_l2_prep": "l2_prep_kernel", "ms_l2_events": "l2_events_kernel",
                    "ms_l2_slide": "l2_slide_kernel", "ms_cgi": "cgi_best_kernel"}
    roofline = {"bound": "hbm", "kernel": kernel_names[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic(kernel_names[top], a),
                "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                "algorithmic_bytes_per_launch": alg[top], "ms_per_launch": top_ms,
                "share_of_step": top_ms / (dev_ms / a.steps)}
    stage_roofline = {k: {"ms": per_step.get(k, 0.0), "alg_bytes": alg[k],
                          "frac": (alg[k] / (per_step[k] * 1e-3) / 1e9 / peak) if per_step.get(k, 0) > 0 else None} for k in alg}

    # ---- CPU baseline + full-size parity on the sampled pairs -----------------------------------
    cpu = None
    parity = None
    if sample_refs:
        cb, ohits, ids = cpu_arm(a, query, sample_refs, 1, 0)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        mine = {h.name: (h.matches, h.fragments, np.float32(h.identity)) for h in hits}
        ok = 0
        for h in ohits:
            gid = ids[int(h["ref_genome"])]
            ok += int(mine.get(gid) == (int(h["matches"]), int(h["fragments"]), np.float32(h["identity"])))
        parity = {"pairs_checked": len(ohits), "identical": ok,
                  "what": "hits of the sampled references: matches, fragments and identity bit-exact vs the CPU " + cb["kind"]}
        assert ok == len(ohits), "GPU hits differ from the CPU reference on the sampled pairs"

    ms_per_step = dev_ms / a.steps
    jobs = 1 if by_refs else world          # whole-job units per step: one query vs all references, or one query per GPU
    line = {
        "metric": METRIC, "value": jobs * pairs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 1), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if by_refs else "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "refs": a.refs, "length": a.length, "fragment_length": FRAG, "k": 16,
                   "window": mapper.window_size,
                   "parallelism": ("reference genomes sharded over the GPUs (%s per rank), same query on every GPU, NCCL all-gather "
                                   "of the hit rows inside the step" % [offsets[r + 1] - offsets[r] for r in range(world)]) if by_refs
                                  else "replicated index, one query stream per GPU",
                   "l2_policy": "inputs larger than L2: the index is %.1f GB, every step streams it" % (n_min * 50 / 1e9)},
        "fragments_per_s": jobs * frags / (ms_per_step * 1e-3),
        "clocks": clocks,
        "e2e": {"value": jobs * pairs * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // a.steps,
                "d2h_bytes_per_step": d2h // a.steps, "ms_per_step": e2e_s / a.steps * 1e3,
                "fragments_per_s": jobs * frags * a.steps / e2e_s},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stages": stage_roofline,
        "counters": {k: inf[k] for k in ("fragments", "sketch_sum", "seeds", "candidates", "scanned", "events", "mappings",
                                         "l2_fallback", "l1_sorted_fragments", "events_replayed")},
        "hits": len(hits),
        "parity": parity,
        "index_build": {"sketch_s": t_sketch, "index_s": t_index, "minimizers": n_min,
                        "sketch_mbp_per_s": a.refs * a.length / 1e6 / t_sketch},
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
