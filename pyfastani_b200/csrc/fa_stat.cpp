// fa_stat.cpp -- see fa_stat.h.  Float/double evaluation order follows the reference
// expression by expression so results are bit-identical:
// everything is double except where the reference stores into a float.
#include "fa_stat.h"

#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

namespace fa {

float j2md(float j, int k)
{
    if (j == 0) return 1.0f;
    if (j == 1) return 0.0f;
    float one_plus_j = 1 + j;                                   // int + float -> float in the reference
    return (float)((-1.0 / k) * std::log(2.0 * j / one_plus_j));
}

float md2j(float d, int k)
{
    float kd = k * d;                                           // int * float -> float
    return (float)(1.0 / (2.0 * std::exp((double)kd) - 1.0));
}

namespace {
// pmf(mode) and the odds ratio; every other term follows by the two-term recurrence.
struct Binom {
    int n, mode; double r, pm;
    Binom(int n_, double p) : n(n_)
    {
        mode = (int)std::floor((n + 1) * p);
        if (mode > n) mode = n;
        r = p / (1 - p);
        pm = std::exp(std::lgamma(n + 1.0) - std::lgamma(mode + 1.0) - std::lgamma(n - mode + 1.0)
                      + mode * std::log(p) + (n - mode) * std::log1p(-p));
    }
    double up(double t, int i) const { return t * r * (double)(n - i) / (double)(i + 1); }     // pmf(i) -> pmf(i+1)
    double down(double t, int i) const { return t * (double)i / ((double)(n - i + 1) * r); }   // pmf(i) -> pmf(i-1)
};
}  // namespace

// P[Bin(n,p) > x]
double binom_sf(int n, double p, int x)
{
    if (x < 0) return 1.0;
    if (x >= n || p <= 0) return 0.0;
    if (p >= 1) return 1.0;
    Binom b(n, p);
    double sum = 0, t = b.pm;
    if (x + 1 > b.mode) {
        for (int i = b.mode; i < x + 1 && t > 0; i++) t = b.up(t, i);
        for (int i = x + 1; i <= n && t > 0; i++) { sum += t; t = b.up(t, i); }
        return sum;
    }
    for (int i = b.mode; i <= n && t > 0; i++) { sum += t; t = b.up(t, i); }
    t = b.pm;
    for (int i = b.mode; i > x + 1 && t > 0; i--) { t = b.down(t, i); sum += t; }
    return sum;
}

// Smallest X with P[Bin(n,p) > X] <= q: what Boost returns for the complemented quantile
// under its default integer_round_outwards policy when q < 0.5.
int binom_quantile_upper(int n, double p, double q)
{
    if (p <= 0) return 0;
    if (p >= 1) return n;
    Binom b(n, p);
    int top = b.mode;
    double t = b.pm;
    while (top < n) {                      // climb until the terms are negligible
        double nt = b.up(t, top);
        if (nt < b.pm * 1e-40) break;
        t = nt; top++;
    }
    double tail = 0;                       // tail(X) = sum_{i > X} pmf(i), accumulated smallest-first
    int X = top;
    while (X > 0) {
        double with_x = tail + t;          // tail(X - 1)
        if (with_x > q) break;
        tail = with_x;
        t = b.down(t, X);
        X--;
    }
    return X;
}

float md_lower_bound(float d, int s, int k, float ci)
{
    float q2 = (float)((1.0 - ci) / 2);
    int x = binom_quantile_upper(s, (double)md2j(d, k), (double)q2);
    float jaccard = (float)x / s;
    return j2md(jaccard, k);
}

static int minimum_hits(int s, int k, float pid)
{
    float mash_dist = (float)(1.0 - pid / 100.0);
    float jaccard = md2j(mash_dist, k);
    return (int)std::ceil(1.0 * s * jaccard);
}

static bool upper_bound_passes(int x, int s, int k, float pid)
{
    float jaccard = (float)(1.0 * x / s);
    float d = j2md(jaccard, k);
    float d_lower = md_lower_bound(d, s, k, 0.9f);
    float id_upper = (float)(100.0 * (1.0 - d_lower));
    return id_upper >= pid;
}

int minimum_hits_relaxed(int s, int k, float pid)
{
    int first = minimum_hits(s, k, pid), relaxed = first;
    for (int i = first; i >= 0; i--) {
        if (upper_bound_passes(i, s, k, pid)) relaxed = i; else break;
    }
    return relaxed;
}

bool l2_pass(int shared, int s, int k, float pid, float *identity)
{
    float mash = j2md((float)(1.0 * shared / s), k);
    float lower = md_lower_bound(mash, s, k, 0.9f);
    float nuc = 100 * (1 - mash);                               // float arithmetic, computeMap.hpp:376
    float upper = 100 * (1 - lower);
    if (identity) *identity = nuc;
    return upper >= pid;
}

static double estimate_pvalue(int s, int k, int alphabet, float pid, int len_query, uint64_t len_ref)
{
    double kmer_space = std::pow((double)alphabet, (double)k);
    double px = 1. / (1. + kmer_space / len_query), py = px;
    double r = px * py / (px + py - px * py);
    int x = minimum_hits_relaxed(s, k, pid);
    double cdfc = (x == 0) ? 1.0 : binom_sf(s, r, x - 1);
    return len_ref * cdfc;
}

int recommended_window(double p_value, int k, int alphabet, float pid, int frag_len, uint64_t ref_size)
{
    int optimal = 0;
    bool found = false;
    auto try_size = [&](int e) {
        if (estimate_pvalue(e, k, alphabet, pid, frag_len, ref_size) <= p_value) { optimal = e; found = true; }
    };
    for (int e : {1, 2, 5}) { if (!found) try_size(e); }
    for (int e = 10; e < frag_len && !found; e += 10) try_size(e);
    // No candidate passes: the reference reads an uninitialised int here (map_stats.hpp:238,253)
    // and in practice returns 1; do the same, deterministically.
    if (!found) return 1;
    int w = (int)(2.0 * frag_len / optimal);
    if (w < 1) w = 1;
    return w < frag_len ? w : frag_len;
}

const StatTable &stat_table(int k, float pid, int s_max)
{
    static std::mutex mtx;
    static std::map<std::tuple<int, float, int>, StatTable> cache;
    std::lock_guard<std::mutex> g(mtx);
    auto key = std::make_tuple(k, pid, s_max);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    StatTable &t = cache[key];
    t.s_max = s_max;
    t.min_hits.assign(s_max + 1, 1);
    t.min_shared.assign(s_max + 1, 0);
    t.id_off.assign(s_max + 2, 0);
    for (int s = 1; s <= s_max; s++) t.id_off[s + 1] = t.id_off[s] + (uint32_t)(s + 1);
    t.identity.assign(t.id_off[s_max + 1], 0.0f);
    // Rows are independent and a row costs O(s) binomial terms per probe, so the table (2 962 rows for the default
    // parameters, 150 ms on one core -- four times the GPU part of an index build) is filled by all host threads, rows
    // dealt round-robin.
    std::atomic<int> irregular{0}, irregular_l2{0};
    auto fill_rows = [&](int first, int stride) {
    for (int s = first; s <= s_max; s += stride) {
        // estimateMinimumHitsRelaxed walks down from m0 while the bound passes (map_stats.hpp:152-165).  The bound is
        // monotone in x (the binomial quantile is monotone in its success probability), so the first failure of that walk
        // is found by bisection -- and the neighbourhood of the answer is then checked against the walk's own rule, so a
        // float-rounding wobble cannot change a result silently: the row falls back to the reference's linear walk.
        int m0 = minimum_hits(s, k, pid), mh = m0;
        if (m0 > s) m0 = mh = s + 1;       // jaccard cut-off above 1 cannot happen for pid in [0,100]
        if (m0 <= s && upper_bound_passes(m0, s, k, pid)) {
            int a = 0, b = m0;
            while (a < b) {
                int mid = (a + b) / 2;
                if (upper_bound_passes(mid, s, k, pid)) b = mid; else a = mid + 1;
            }
            mh = a;
            bool regular = mh == 0 || !upper_bound_passes(mh - 1, s, k, pid);
            for (int x = mh; x <= std::min(m0, mh + 4) && regular; x++) regular = upper_bound_passes(x, s, k, pid);
            if (!regular) { mh = minimum_hits_relaxed(s, k, pid); irregular++; }
        }
        t.min_hits[s] = mh < 1 ? 1 : mh;
        float *row = &t.identity[t.id_off[s]];
        for (int x = 0; x <= s; x++) {
            float mash = j2md((float)(1.0 * x / s), k);
            row[x] = 100 * (1 - mash);
        }
        // The filter of computeMap.hpp:380 is evaluated per candidate, x by x.  Its CI upper bound is non-decreasing in x,
        // so it is x >= min_shared[s]; the threshold is bisected and its neighbourhood checked (tests/test_abi.py checks
        // whole rows against fa_stat_l2).  A row that is not a clean step is scanned in full: the threshold becomes the
        // smallest x from which everything passes, and passing values below it are counted in `irregular` -- the index
        // build refuses such parameters instead of mapping with a filter that differs from the reference's.
        int lo = 0, hi = s + 1;
        while (lo < hi) {
            int mid = (lo + hi) / 2;
            if (l2_pass(mid, s, k, pid, nullptr)) hi = mid; else lo = mid + 1;
        }
        bool step = true;
        for (int x = std::max(0, lo - 4); x < lo && step; x++) step = !l2_pass(x, s, k, pid, nullptr);
        for (int x = lo; x <= std::min(s, lo + 4) && step; x++) step = l2_pass(x, s, k, pid, nullptr);
        if (!step) {
            int thr = s + 1;
            for (int x = s; x >= 0 && l2_pass(x, s, k, pid, nullptr); x--) thr = x;
            for (int x = 0; x < thr; x++) if (l2_pass(x, s, k, pid, nullptr)) { irregular_l2++; break; }
            lo = thr;
        }
        t.min_shared[s] = lo;
    }
    };
    {
        int workers = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        if (s_max < 64) workers = 1;
        std::vector<std::thread> pool;
        for (int w = 1; w < workers; w++) pool.emplace_back(fill_rows, 1 + w, workers);
        fill_rows(1, workers);
        for (auto &th : pool) th.join();
    }
    t.irregular += irregular.load();
    t.irregular_l2 += irregular_l2.load();
    return t;
}

}  // namespace fa
