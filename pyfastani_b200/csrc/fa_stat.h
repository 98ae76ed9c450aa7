// fa_stat.h -- host-side statistics of the mapping path (functions of (s, x, k, pid) only).
//
// The reference recomputes these per fragment / per candidate on the CPU
// (FA/map/include/map_stats.hpp:44-256, called from computeMap.hpp:371-377 and
// pyx:951); here they are tabulated once per index and uploaded, so the device
// never evaluates log/exp and the float results are bit-identical to the
// reference's.
#pragma once
#include <cstdint>
#include <vector>

namespace fa {

float j2md(float j, int k);                                   // map_stats.hpp:44-54
float md2j(float d, int k);                                   // map_stats.hpp:62-66
int   binom_quantile_upper(int n, double p, double q);        // Boost quantile(complement(binomial)) stand-in, :88
double binom_sf(int n, double p, int x);                      // Boost cdf(complement(binomial)) stand-in, :204
float md_lower_bound(float d, int s, int k, float ci);        // map_stats.hpp:79-111
int   minimum_hits_relaxed(int s, int k, float pid);          // map_stats.hpp:142-167
bool  l2_pass(int shared, int s, int k, float pid, float *identity);   // computeMap.hpp:371-380
int   recommended_window(double p_value, int k, int alphabet, float pid, int frag_len, uint64_t ref_size);  // :226-256

// Tables for sketch sizes 1..s_max (row s at id_off[s]; identity[id_off[s] + x], x in [0, s]).
struct StatTable {
    int s_max = 0;
    std::vector<int32_t>  min_hits;     // max(1, estimateMinimumHitsRelaxed(s)); computeMap.hpp:312-313
    std::vector<int32_t>  min_shared;   // smallest x whose CI upper bound passes; s+1 if none
    std::vector<uint32_t> id_off;       // s_max + 2 entries
    std::vector<float>    identity;     // nucIdentity(x, s), computeMap.hpp:376
    int irregular = 0;                  // rows whose minHits bisection disagreed with the reference's walk (the walk's value is stored)
    int irregular_l2 = 0;               // rows whose L2 filter is not a step in x: no single threshold reproduces the reference
};
// Cached per (k, pid, s_max); thread-safe.
const StatTable &stat_table(int k, float pid, int s_max);

}  // namespace fa
