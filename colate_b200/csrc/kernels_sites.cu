// Stage i kernels (parse_tmptmp, coal.cpp:2071-2321) on structure-of-arrays in HBM:
//   k_join      per genome: merge-join of the .colate.in records onto the site axis
//   k_cand/k_ok per pair: row filter + the two stream lookups incl. the sequential reader's
//               look-ahead rule as bitmap passes
//   k_*rank*    used-row rank = offset into the reference's generator stream
//   k_compact   used rows -> dense records in rank order, genomic block index
//   k_sample    THE per-mutation kernel: 100 Monte-Carlo age draws per used row from the
//               uniform stream in HBM, exact bin index, warp-aggregated histogram updates in
//               per-warp shared-memory histograms (no atomics), one partial per tile
//   k_reduce    fixed-order sum of the tile partials of each genomic block
#include "device.cuh"

namespace colate {

constexpr int TILE_SITES = 1024;  // used rows per sampling tile
constexpr int SAMPLE_WARPS = 8;
constexpr int SCAN_ITEMS = 8;     // bitmap words per thread in the rank scan
constexpr int SCAN_THREADS = 256;

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int chr_of(const int64_t* __restrict__ site_off, int n_chr, int64_t m)
{
  int lo = 0, hi = n_chr;  // largest c with site_off[c] <= m
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (site_off[mid] <= m) lo = mid; else hi = mid;
  }
  return lo;
}

// largest p in [lo, m) whose bit is set, else -1
__device__ __forceinline__ int64_t prev_set(const uint32_t* __restrict__ bits, int64_t m, int64_t lo)
{
  int64_t w = m >> 5;
  uint32_t x = (m & 31) ? (bits[w] & ((1u << (m & 31)) - 1u)) : 0u;
  for (;;) {
    if (x) {
      int64_t p = (w << 5) + 31 - __clz(x);
      return p >= lo ? p : -1;
    }
    if ((w << 5) <= lo) return -1;
    w--;
    x = bits[w];
  }
}

__device__ __forceinline__ int64_t rank_of(const uint32_t* __restrict__ use, const uint32_t* __restrict__ word_rank, int64_t m)
{
  int64_t w = m >> 5;
  uint32_t r = word_rank[w];
  if (m & 31) r += __popc(use[w] & ((1u << (m & 31)) - 1u));
  return r;
}

// ------------------------------------------------------------------------------------------
__global__ void k_join(int64_t n_site, int n_chr, const int64_t* __restrict__ site_off,
                       const int32_t* __restrict__ pos, const uint32_t* __restrict__ meta,
                       const int64_t* __restrict__ chr_first, const int64_t* __restrict__ chr_end,
                       const int32_t* __restrict__ bp, const int32_t* __restrict__ aaf,
                       const int32_t* __restrict__ daf, const uint16_t* __restrict__ alleles,
                       int32_t* __restrict__ j_aaf, int32_t* __restrict__ j_daf,
                       int32_t* __restrict__ j_prevbp, uint8_t* __restrict__ j_flag)
{
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < n_site; m += (int64_t)gridDim.x * blockDim.x) {
    int32_t a = 0, d = 0, pb = -1;
    uint8_t fl = 0;
    uint32_t mt = meta[m];
    if (mt & 1u) {
      int c = chr_of(site_off, n_chr, m);
      int64_t first = chr_first[c], end = chr_end[c];
      if (first >= 0) {
        int32_t p = pos[m];
        int64_t lo = first, hi = end;
        while (lo < hi) {
          int64_t mid = (lo + hi) >> 1;
          if (bp[mid] < p) lo = mid + 1; else hi = mid;
        }
        if (lo < end && bp[lo] == p) {
          fl = 1;
          a = aaf[lo];
          d = daf[lo];
          pb = lo > first ? bp[lo - 1] : -1;
          uint32_t al = alleles[lo];
          if ((al & 0xffu) == ((mt >> 8) & 0xffu) && (al >> 8) == ((mt >> 16) & 0xffu)) fl |= 2;
        }
      }
    }
    j_aaf[m] = a; j_daf[m] = d; j_prevbp[m] = pb; j_flag[m] = fl;
  }
}

// row filter x masks -> candidate bitmap for the reference stream (coal.cpp:2150-2181)
__global__ void k_cand(int64_t n_site, const uint32_t* __restrict__ meta, const uint32_t* __restrict__ tmask,
                       const uint32_t* __restrict__ rmask, uint32_t* __restrict__ cand)
{
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool c = false;
  if (m < n_site) {
    c = meta[m] & 1u;
    if (tmask) c = c && ((tmask[m >> 5] >> (m & 31)) & 1u);
    if (rmask) c = c && ((rmask[m >> 5] >> (m & 31)) & 1u);
  }
  uint32_t b = __ballot_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && m < n_site) cand[m >> 5] = b;
}

// one stream lookup for every candidate row (coal.cpp:2181-2199 / 2201-2219): the row keeps
// `use` iff a record sits at its position with the same alleles, that record had not already
// been pulled in by the look-ahead of an earlier candidate (or by the chromosome seek), and
// DAF_ref != 0 (reference stream) / AAF+DAF != 0 (target stream).
template <bool IS_REF>
__global__ void k_ok(int64_t n_site, int n_chr, const int64_t* __restrict__ site_off, const int32_t* __restrict__ pos,
                     const uint32_t* __restrict__ in_bits, const int32_t* __restrict__ j_aaf,
                     const int32_t* __restrict__ j_daf, const int32_t* __restrict__ j_prevbp,
                     const uint8_t* __restrict__ j_flag, uint32_t* __restrict__ out_bits)
{
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (m < n_site && ((in_bits[m >> 5] >> (m & 31)) & 1u)) {
    if ((j_flag[m] & 3) == 3) {
      bool cnt = IS_REF ? (j_daf[m] != 0) : ((j_aaf[m] + j_daf[m]) != 0);
      if (cnt) {
        int64_t lo = site_off[chr_of(site_off, n_chr, m)];
        int64_t p = prev_set(in_bits, m, lo);
        int32_t prev_cand_pos = p >= 0 ? pos[p] : 0;
        ok = prev_cand_pos <= j_prevbp[m];
      }
    }
  }
  uint32_t b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m < n_site) out_bits[m >> 5] = b;
}

// ---- rank of every bitmap word (exclusive prefix popcount) --------------------------------
__global__ void k_popc_blocksum(const uint32_t* __restrict__ words, int64_t n_words, uint32_t* __restrict__ sums)
{
  __shared__ uint32_t s[SCAN_THREADS / 32];
  int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  uint32_t v = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n_words) v += __popc(words[base + i]);
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < SCAN_THREADS / 32; i++) t += s[i];
    sums[blockIdx.x] = t;
  }
}

__global__ void k_scan_sums(uint32_t* __restrict__ sums, int n, uint32_t* __restrict__ total)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {  // n is a few hundred
    uint32_t run = 0;
    for (int i = 0; i < n; i++) { uint32_t v = sums[i]; sums[i] = run; run += v; }
    *total = run;
  }
}

__global__ void k_word_rank(const uint32_t* __restrict__ words, int64_t n_words, const uint32_t* __restrict__ block_off,
                            const uint32_t* __restrict__ total, uint32_t* __restrict__ word_rank)
{
  __shared__ uint32_t s[SCAN_THREADS / 32];
  int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  uint32_t loc[SCAN_ITEMS];
  uint32_t v = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) { loc[i] = v; if (base + i < n_words) v += __popc(words[base + i]); }
  uint32_t incl = v;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s[wid] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int i = 0; i < wid; i++) woff += s[i];
  uint32_t excl = block_off[blockIdx.x] + woff + incl - v;
  for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n_words) word_rank[base + i] = excl + loc[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) word_rank[n_words] = *total;
}

// per chromosome: used rows, genomic blocks (coal.cpp:2227-2234, 2306-2310), block base
__global__ void k_chr(int n_chr, const int64_t* __restrict__ site_off, const int32_t* __restrict__ pos,
                      const uint32_t* __restrict__ use, const uint32_t* __restrict__ word_rank,
                      int64_t* __restrict__ chr_used, int32_t* __restrict__ chr_blocks, int32_t* __restrict__ chr_block_base,
                      int64_t* __restrict__ misc)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int base = 0;
  int64_t tot = 0;
  for (int c = 0; c < n_chr; c++) {
    int64_t lo = site_off[c], hi = site_off[c + 1];
    int64_t used = rank_of(use, word_rank, hi) - rank_of(use, word_rank, lo);
    int64_t p = hi > lo ? prev_set(use, hi, lo) : -1;
    int nb = p >= 0 ? (pos[p] - 1) / COLATE_BLOCK_BASES + 1 : 1;
    chr_used[c] = used;
    chr_blocks[c] = nb;
    chr_block_base[c] = base;
    base += nb;
    tot += used;
  }
  misc[0] = tot;
  misc[1] = base;
}

// used rows -> dense records in rank order
__global__ void k_compact(int64_t n_site, int n_chr, const int64_t* __restrict__ site_off, const int32_t* __restrict__ pos,
                          const float* __restrict__ ab, const float* __restrict__ ae,
                          const uint32_t* __restrict__ use, const uint32_t* __restrict__ word_rank,
                          const int32_t* __restrict__ chr_block_base,
                          const int32_t* __restrict__ t_aaf, const int32_t* __restrict__ t_daf,
                          const int32_t* __restrict__ r_aaf, const int32_t* __restrict__ r_daf,
                          const double* __restrict__ thr10,
                          float* __restrict__ u_ab, float* __restrict__ u_ae, float* __restrict__ u_fd, float* __restrict__ u_fa,
                          int32_t* __restrict__ u_dafr, int32_t* __restrict__ u_nr, int32_t* __restrict__ u_blk,
                          int64_t* __restrict__ misc)
{
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_site) return;
  if (!((use[m >> 5] >> (m & 31)) & 1u)) return;
  int64_t r = rank_of(use, word_rank, m);
  int c = chr_of(site_off, n_chr, m);
  // pseudo-genotype, coal.cpp:2236-2242: float /= double, then round half away
  int32_t dt = t_daf[m], at = t_aaf[m];
  double half_n = (double)(dt + at) / 2.0;
  float fd = __double2float_rn(__ddiv_rn((double)(float)dt, half_n));
  float fa = __double2float_rn(__ddiv_rn((double)(float)at, half_n));
  fd = roundf(fd);
  fa = roundf(fa);
  float b = ab[m];
  if (b < 0.0f) b = 0.0f;  // coal.cpp:2225 with ref_age == 0 (coal.cpp:2075)
  float e = ae[m];
  u_ab[r] = b;
  u_ae[r] = e;
  u_fd[r] = fd;
  u_fa[r] = fa;
  u_dafr[r] = r_daf[m];
  u_nr[r] = r_daf[m] + r_aaf[m];
  u_blk[r] = chr_block_base[c] + (pos[m] - 1) / COLATE_BLOCK_BASES;
  // the reference writes out of bounds / rejection-samples once a bin index reaches 185
  if (__dmul_rn(10.0, (double)e) >= thr10[NBINS]) misc[3] = 1;
}

// rank range of every genomic block and the tile table
__global__ void k_tiles(const int32_t* __restrict__ u_blk, int64_t* __restrict__ misc,
                        int64_t* __restrict__ blk_rank_start, int32_t* __restrict__ tile_start)
{
  int64_t n_used = misc[0];
  int n_blocks = (int)misc[1];
  for (int b = threadIdx.x; b <= n_blocks; b += blockDim.x) {
    int64_t lo = 0, hi = n_used;  // first rank with u_blk >= b
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (u_blk[mid] < b) lo = mid + 1; else hi = mid;
    }
    blk_rank_start[b] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int b = 0; b < n_blocks; b++) {
      tile_start[b] = t;
      t += (int)((blk_rank_start[b + 1] - blk_rank_start[b] + TILE_SITES - 1) / TILE_SITES);
    }
    tile_start[n_blocks] = t;
    misc[2] = t;
  }
}

// ------------------------------------------------------------------------------------------
// uniform_real_distribution<double>(0,1) from two engine words, bit-exact with libstdc++'s
// generate_canonical<double,53>: (x1 + x2*2^32) / 2^64, rounded once, clamped below 1.
__device__ __forceinline__ double u01(uint32_t x1, uint32_t x2)
{
  double d1 = __hiloint2double(0x3F300000, (int)x1) - 0x1p-12;  // x1 * 2^-64, exact
  double d2 = __hiloint2double(0x41300000, (int)x2) - 0x1p20;   // x2 * 2^-32, exact
  double u = __dadd_rn(d2, d1);
  return u >= 1.0 ? 0x1.fffffffffffffp-1 : u;
}

// max(0,(int)round(log(x10)*10)+1) (coal.cpp:2253/2265/2284) without evaluating log in fp64:
// a float estimate picks the bin, the exact host-computed thresholds settle it.
__device__ __forceinline__ int bin_of_x10(double x10, const double* thr)
{
  float lf = __log2f((float)x10) * 6.931471805599453f;
  int k = __float2int_rn(lf);
  k = max(0, min(k + 1, NBINS));
  while (k < NBINS && x10 >= thr[k + 1]) k++;
  while (k > 0 && x10 < thr[k]) k--;
  return k;
}

struct WarpHist {
  double hS[NBINS], hN[NBINS], eS[NBINS], eN[NBINS];
  uint32_t cS[NBINS], cN[NBINS], cE[NBINS];
  uint32_t pad;
};

__global__ void __launch_bounds__(SAMPLE_WARPS * 32)
k_sample(const int64_t* misc_in, const int64_t* __restrict__ blk_rank_start,
         const int32_t* __restrict__ tile_start, const double* __restrict__ thr10_g,
         const float* __restrict__ u_ab, const float* __restrict__ u_ae, const float* __restrict__ u_fd,
         const float* __restrict__ u_fa, const int32_t* __restrict__ u_dafr, const int32_t* __restrict__ u_nr,
         const uint32_t* __restrict__ stream, double* __restrict__ partial_f, uint32_t* __restrict__ partial_n,
         int64_t* misc)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpHist* hist = (WarpHist*)smem_raw;
  double* thr = (double*)(smem_raw + sizeof(WarpHist) * SAMPLE_WARPS);

  const int n_tiles = (int)misc_in[2];
  const int n_blocks = (int)misc_in[1];
  const int tile = blockIdx.x;
  if (tile >= n_tiles) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < NTHR; i += blockDim.x) thr[i] = thr10_g[i];
  {
    uint32_t* z = (uint32_t*)&hist[warp];
    for (int i = lane; i < (int)(sizeof(WarpHist) / 4); i += 32) z[i] = 0;
  }
  // genomic block of this tile: last b with tile_start[b] <= tile
  int lo = 0, hi = n_blocks;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (tile_start[mid] <= tile) lo = mid; else hi = mid;
  }
  const int blk = lo;
  const int64_t r0 = blk_rank_start[blk] + (int64_t)(tile - tile_start[blk]) * TILE_SITES;
  const int64_t r1 = min(r0 + TILE_SITES, blk_rank_start[blk + 1]);
  __syncthreads();

  WarpHist& H = hist[warp];
  bool overflow = false;
  for (int64_t r = r0 + warp; r < r1; r += SAMPLE_WARPS) {
    const float abf = u_ab[r], aef = u_ae[r], fd = u_fd[r], fa = u_fa[r];
    const int32_t dafr = u_dafr[r], nr = u_nr[r];
    const double abd = (double)abf;
    const double len = __dsub_rn((double)aef, abd);
    const bool emp = abd <= 0.0;  // coal.cpp:2247 with age == 0
    const double num_s = (double)__fmul_rn(fd, __int2float_rn(dafr));
    const double num_n = (double)__fmul_rn(fa, __int2float_rn(dafr));
    const double den = __dmul_rn((double)nr, 100.0);
    const double wS = __ddiv_rn(num_s, den), wN = __ddiv_rn(num_n, den);

    if (emp && lane == 0) {  // coal.cpp:2250-2256
      int b2 = bin_of_x10((double)__fmul_rn(10.0f, aef), thr);
      if (b2 < NBINS) {
        H.eS[b2] += __ddiv_rn(num_s, (double)nr);
        H.eN[b2] += __ddiv_rn(num_n, (double)nr);
        H.cE[b2] += 1;
      }
    }
    // 200 engine words of this row: 50 x 16 B; lane l takes chunks l and 32+l
    const uint4* sp = (const uint4*)(stream + 200 * r);
    uint4 q0 = __ldg(sp + lane);
    uint4 q1 = make_uint4(0, 0, 0, 0);
    const bool has1 = lane < 18;
    if (has1) q1 = __ldg(sp + 32 + lane);
    int bins[4];
    {
      double a;
      a = __dadd_rn(__dmul_rn(u01(q0.x, q0.y), len), abd); bins[0] = bin_of_x10(__dmul_rn(10.0, a), thr);
      a = __dadd_rn(__dmul_rn(u01(q0.z, q0.w), len), abd); bins[1] = bin_of_x10(__dmul_rn(10.0, a), thr);
      a = __dadd_rn(__dmul_rn(u01(q1.x, q1.y), len), abd); bins[2] = has1 ? bin_of_x10(__dmul_rn(10.0, a), thr) : 0xffff;
      a = __dadd_rn(__dmul_rn(u01(q1.z, q1.w), len), abd); bins[3] = has1 ? bin_of_x10(__dmul_rn(10.0, a), thr) : 0xffff;
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int b = bins[s];
      const unsigned peers = __match_any_sync(0xffffffffu, b);
      if (b < NBINS) {
        if (lane == __ffs(peers) - 1) {
          const int cnt = __popc(peers);
          const double c = (double)cnt;
          H.hN[b] += __dmul_rn(c, wN);
          H.cN[b] += cnt;
          if (!emp) { H.hS[b] += __dmul_rn(c, wS); H.cS[b] += cnt; }
        }
      } else if (b == NBINS) overflow = true;
      __syncwarp();
    }
  }
  if (overflow) misc[3] = 1;
  __syncthreads();
  // fixed-order sum over the warps of the tile
  double* pf = partial_f + (size_t)tile * 4 * NBINS;
  uint32_t* pn = partial_n + (size_t)tile * 3 * NBINS;
  for (int i = threadIdx.x; i < NBINS; i += blockDim.x) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    for (int w = 0; w < SAMPLE_WARPS; w++) {
      s0 += hist[w].hS[i]; s1 += hist[w].hN[i]; s2 += hist[w].eS[i]; s3 += hist[w].eN[i];
      c0 += hist[w].cS[i]; c1 += hist[w].cN[i]; c2 += hist[w].cE[i];
    }
    pf[i] = s0; pf[NBINS + i] = s1; pf[2 * NBINS + i] = s2; pf[3 * NBINS + i] = s3;
    pn[i] = c0; pn[NBINS + i] = c1; pn[2 * NBINS + i] = c2;
  }
}

__global__ void k_reduce(const int32_t* __restrict__ tile_start, const double* __restrict__ partial_f,
                         const uint32_t* __restrict__ partial_n, double* __restrict__ out_f, int64_t* __restrict__ out_n)
{
  const int blk = blockIdx.x;
  const int t0 = tile_start[blk], t1 = tile_start[blk + 1];
  for (int i = threadIdx.x; i < 4 * NBINS; i += blockDim.x) {
    double s = 0;
    for (int t = t0; t < t1; t++) s += partial_f[(size_t)t * 4 * NBINS + i];
    out_f[(size_t)blk * 4 * NBINS + i] = s;
  }
  for (int i = threadIdx.x; i < 3 * NBINS; i += blockDim.x) {
    int64_t s = 0;
    for (int t = t0; t < t1; t++) s += partial_n[(size_t)t * 3 * NBINS + i];
    out_n[(size_t)blk * 3 * NBINS + i] = s;
  }
}

// ------------------------------------------------------------------------------------------
static inline int grid_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

int run_join(colate_handle* h, int slot)
{
  GenomeDev& g = h->genomes[slot];
  if (g.joined) return 0;
  const int64_t n = h->n_site;
  CK(g.j_aaf.ensure(n * 4 + 4)); CK(g.j_daf.ensure(n * 4 + 4)); CK(g.j_prevbp.ensure(n * 4 + 4)); CK(g.j_flag.ensure(n + 4));
  if (n > 0) {
    int grid = (int)std::min<int64_t>(grid_for(n, 256), 148 * 16);
    k_join<<<grid, 256, 0, h->stream>>>(n, h->n_chr, h->site_off.as<int64_t>(), h->pos.as<int32_t>(), h->meta.as<uint32_t>(),
                                        g.chr_first.as<int64_t>(), g.chr_end.as<int64_t>(), g.bp.as<int32_t>(),
                                        g.aaf.as<int32_t>(), g.daf.as<int32_t>(), g.alleles.as<uint16_t>(),
                                        g.j_aaf.as<int32_t>(), g.j_daf.as<int32_t>(), g.j_prevbp.as<int32_t>(), g.j_flag.as<uint8_t>());
    CK(cudaGetLastError());
    h->launches += 1;
  }
  g.joined = true;
  return 0;
}

int run_flags(colate_handle* h, int tslot, int rslot)
{
  const int64_t n = h->n_site;
  const int64_t nw = (n + 31) / 32;
  GenomeDev& T = h->genomes[tslot];
  GenomeDev& R = h->genomes[rslot];
  CK(h->candR.ensure(nw * 4 + 8)); CK(h->candT.ensure(nw * 4 + 8)); CK(h->use.ensure(nw * 4 + 8));
  CK(h->word_rank.ensure((nw + 1) * 4 + 8));
  const int nsb = std::max(1, grid_for(nw, SCAN_THREADS * SCAN_ITEMS));
  CK(h->scan_tmp.ensure((size_t)nsb * 4 + 16));
  CK(h->chr_used.ensure(h->n_chr * 8 + 8)); CK(h->chr_blocks.ensure(h->n_chr * 4 + 8)); CK(h->chr_block_base.ensure(h->n_chr * 4 + 8));
  CK(h->misc.ensure(64));
  CK(cudaMemsetAsync(h->misc.p, 0, 64, h->stream));
  cudaStream_t s = h->stream;
  uint32_t* total = (uint32_t*)((char*)h->misc.p + 56);
  if (n > 0) {
    int grid = grid_for(n, 256);
    k_cand<<<grid, 256, 0, s>>>(n, h->meta.as<uint32_t>(), T.has_mask ? T.mask_bits.as<uint32_t>() : nullptr,
                                R.has_mask ? R.mask_bits.as<uint32_t>() : nullptr, h->candR.as<uint32_t>());
    k_ok<true><<<grid, 256, 0, s>>>(n, h->n_chr, h->site_off.as<int64_t>(), h->pos.as<int32_t>(), h->candR.as<uint32_t>(),
                                    R.j_aaf.as<int32_t>(), R.j_daf.as<int32_t>(), R.j_prevbp.as<int32_t>(), R.j_flag.as<uint8_t>(),
                                    h->candT.as<uint32_t>());
    k_ok<false><<<grid, 256, 0, s>>>(n, h->n_chr, h->site_off.as<int64_t>(), h->pos.as<int32_t>(), h->candT.as<uint32_t>(),
                                     T.j_aaf.as<int32_t>(), T.j_daf.as<int32_t>(), T.j_prevbp.as<int32_t>(), T.j_flag.as<uint8_t>(),
                                     h->use.as<uint32_t>());
    k_popc_blocksum<<<nsb, SCAN_THREADS, 0, s>>>(h->use.as<uint32_t>(), nw, h->scan_tmp.as<uint32_t>());
    k_scan_sums<<<1, 32, 0, s>>>(h->scan_tmp.as<uint32_t>(), nsb, total);
    k_word_rank<<<nsb, SCAN_THREADS, 0, s>>>(h->use.as<uint32_t>(), nw, h->scan_tmp.as<uint32_t>(), total, h->word_rank.as<uint32_t>());
    h->launches += 6;
  } else {
    CK(cudaMemsetAsync(h->word_rank.p, 0, 8, s));
  }
  k_chr<<<1, 32, 0, s>>>(h->n_chr, h->site_off.as<int64_t>(), h->pos.as<int32_t>(), h->use.as<uint32_t>(), h->word_rank.as<uint32_t>(),
                         h->chr_used.as<int64_t>(), h->chr_blocks.as<int32_t>(), h->chr_block_base.as<int32_t>(), h->misc.as<int64_t>());
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// compaction + tile table + sampling + per-block reduction; needs h->n_used / n_blocks_local
int run_sample(colate_handle* h, const uint32_t* stream_local, int)
{
  const int64_t n = h->n_site, nu = h->n_used;
  const int nb = h->n_blocks_local;
  GenomeDev& T = h->genomes[h->tgt_slot];
  GenomeDev& R = h->genomes[h->ref_slot];
  cudaStream_t s = h->stream;
  const size_t un = (size_t)std::max<int64_t>(nu, 1);
  CK(h->u_ab.ensure(un * 4)); CK(h->u_ae.ensure(un * 4)); CK(h->u_fd.ensure(un * 4)); CK(h->u_fa.ensure(un * 4));
  CK(h->u_dafr.ensure(un * 4)); CK(h->u_nr.ensure(un * 4)); CK(h->u_blk.ensure(un * 4));
  CK(h->blk_rank_start.ensure((MAX_BLOCKS + 2) * 8)); CK(h->tile_start.ensure((MAX_BLOCKS + 2) * 4));
  const int max_tiles = (int)(nu / TILE_SITES) + nb + 1;
  CK(h->partial_f.ensure((size_t)max_tiles * 4 * NBINS * 8)); CK(h->partial_n.ensure((size_t)max_tiles * 3 * NBINS * 4));
  CK(h->out_f.ensure((size_t)MAX_BLOCKS * 4 * NBINS * 8)); CK(h->out_n.ensure((size_t)MAX_BLOCKS * 3 * NBINS * 8));
  CK(cudaEventRecord(h->ev[2], s));
  h->launches += (n > 0 ? 1 : 0) + 2 + (nb > 0 ? 1 : 0);
  if (n > 0)
    k_compact<<<grid_for(n, 256), 256, 0, s>>>(n, h->n_chr, h->site_off.as<int64_t>(), h->pos.as<int32_t>(), h->ab.as<float>(),
                                               h->ae.as<float>(), h->use.as<uint32_t>(), h->word_rank.as<uint32_t>(),
                                               h->chr_block_base.as<int32_t>(), T.j_aaf.as<int32_t>(), T.j_daf.as<int32_t>(),
                                               R.j_aaf.as<int32_t>(), R.j_daf.as<int32_t>(), h->thr10.as<double>(),
                                               h->u_ab.as<float>(), h->u_ae.as<float>(), h->u_fd.as<float>(), h->u_fa.as<float>(),
                                               h->u_dafr.as<int32_t>(), h->u_nr.as<int32_t>(), h->u_blk.as<int32_t>(), h->misc.as<int64_t>());
  k_tiles<<<1, 512, 0, s>>>(h->u_blk.as<int32_t>(), h->misc.as<int64_t>(), h->blk_rank_start.as<int64_t>(), h->tile_start.as<int32_t>());
  CK(cudaEventRecord(h->ev[3], s));
  const size_t smem = sizeof(WarpHist) * SAMPLE_WARPS + NTHR * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_sample<<<max_tiles, SAMPLE_WARPS * 32, smem, s>>>(h->misc.as<int64_t>(), h->blk_rank_start.as<int64_t>(), h->tile_start.as<int32_t>(),
                                                      h->thr10.as<double>(), h->u_ab.as<float>(), h->u_ae.as<float>(), h->u_fd.as<float>(),
                                                      h->u_fa.as<float>(), h->u_dafr.as<int32_t>(), h->u_nr.as<int32_t>(), stream_local,
                                                      h->partial_f.as<double>(), h->partial_n.as<uint32_t>(), h->misc.as<int64_t>());
  CK(cudaEventRecord(h->ev[4], s));
  if (nb > 0)
    k_reduce<<<nb, 256, 0, s>>>(h->tile_start.as<int32_t>(), h->partial_f.as<double>(), h->partial_n.as<uint32_t>(),
                                h->out_f.as<double>(), h->out_n.as<int64_t>());
  CK(cudaEventRecord(h->ev[5], s));
  CK(cudaGetLastError());
  return 0;
}

}  // namespace colate
