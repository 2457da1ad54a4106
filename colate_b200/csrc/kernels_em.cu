// Stage ii (block bootstrap + F redistribution, coal.cpp:3344-3451) and stage iii (EM on the
// 185-bin age histograms: coal.cpp:3675-3827 driving coal_EM, coal_EM.cpp:5-468) kernels.
//
// Everything here is IEEE fp64 in the reference's operation order (the file is compiled with
// --fmad=false; products and sums are never contracted).  exp/log/log1p are CUDA's, which
// differ from glibc's in the last ulp: stage ii is bit-exact with the reference, stage iii
// agrees to ~1e-13 relative per iteration (north-star tolerance 1e-9 on the rates).
#include "device.cuh"

namespace colate {

constexpr int EM_THREADS = 384;          // warps 0-5: shared tasks, warps 6-11: not-shared tasks
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int EM_HALF = EM_THREADS / 2;  // 192 >= NBINS

// ---- stage ii ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bootstrap(int num_blocks, const int32_t* __restrict__ weights, const double* __restrict__ blk,
            double age, const double* __restrict__ age_bin, double* __restrict__ counts)
{
  __shared__ double S[NBINS], N[NBINS], SE[NBINS], NE[NBINS], F[NBINS];
  const int r = blockIdx.x;
  const int32_t* w = weights + (size_t)r * num_blocks;
  for (int b = threadIdx.x; b < NBINS; b += blockDim.x) {
    double s = 0.0, n = 0.0, se = 0.0, ne = 0.0;
    for (int j = 0; j < num_blocks; j++) {  // coal.cpp:3358-3390, block order kept
      const double bw = (double)w[j];
      if (bw > 0.0) {
        const double* v = blk + (size_t)j * 4 * NBINS + b;
        s = __dadd_rn(s, __dmul_rn(bw, v[0]));
        n = __dadd_rn(n, __dmul_rn(bw, v[NBINS]));
        se = __dadd_rn(se, __dmul_rn(bw, v[2 * NBINS]));
        ne = __dadd_rn(ne, __dmul_rn(bw, v[3 * NBINS]));
      }
    }
    S[b] = s; N[b] = n; SE[b] = se; NE[b] = ne; F[b] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // serial sums in bin order, coal.cpp:3392-3441
    int bin = 0;
    while (bin < NBINS - 1 && age_bin[bin] <= age) bin++;
    const int bin_start = bin;
    double lower_age = age_bin[bin_start - 1];
    double fcount = 0.0;
    for (bin = bin_start; bin < NBINS; bin++) {
      fcount = __dadd_rn(fcount, SE[bin]);
      if (SE[bin] > 0) F[bin] = __ddiv_rn(SE[bin], __dadd_rn(SE[bin], NE[bin]));
    }
    for (bin = bin_start; bin < NBINS; bin++) {
      F[bin - 1] = __dmul_rn(F[bin - 1], __dsub_rn(age_bin[bin], lower_age));
      lower_age = age_bin[bin];
    }
    double normf = 0.0;
    for (bin = 0; bin < NBINS; bin++) normf = __dadd_rn(normf, F[bin]);
    for (bin = 0; bin < NBINS; bin++) {
      double f = __dmul_rn(__ddiv_rn(F[bin], normf), fcount);
      S[bin] = __dadd_rn(S[bin], (0.0 < f) ? f : 0.0);  // std::max(0.0, f): NaN -> 0.0
    }
  }
  __syncthreads();
  double* out = counts + (size_t)r * 2 * NBINS;
  for (int b = threadIdx.x; b < NBINS; b += blockDim.x) { out[b] = S[b]; out[NBINS + b] = N[b]; }
}

// ---- E-step pieces (coal_EM.cpp) -----------------------------------------------------------
__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ull); }
__device__ __forceinline__ bool bad(double x) { return !(fabs(x) < __longlong_as_double(0x7ff0000000000000ull)); }

// coal_EM::logsumexp, coal_EM.cpp:5-31
__device__ __forceinline__ double lse(double a, double b)
{
  if (bad(a)) return bad(b) ? neg_inf() : b;
  if (bad(b)) return a;
  if (a > b) return a + log1p(exp(b - a));
  return b + log1p(exp(a - b));
}
// coal_EM::logminusexp, coal_EM.cpp:33-58
__device__ __forceinline__ double lme(double a, double b)
{
  if (bad(a)) return neg_inf();
  if (bad(b)) return a;
  if (a < b) return neg_inf();
  return a + log1p(-exp(b - a));
}

struct EmCtx {
  int E;
  const double *ep, *rate, *A, *B, *Lam;
};

// cumulative hazard of the plain epoch grid (coal_EM.cpp:100-103); serial by construction
__device__ void em_cumhaz(int E, const double* ep, const double* rate, double* Lam)
{
  double l = 0.0;
  Lam[0] = 0.0;
  for (int i = 1; i < E; i++) { l = l + rate[i - 1] * (ep[i] - ep[i - 1]); Lam[i] = l; }
}

// coal_EM ctor -> get_AB, coal_EM.cpp:105-149, entry i
__device__ void em_AB(int E, const double* ep, const double* rate, const double* Lam, int i, double* A, double* B)
{
  const double r = rate[i];
  if (i < E - 1) {
    const double tb = ep[i], te = ep[i + 1], inv = 1.0 / r;
    if (r > 0 && te != 0 && te - tb > 0) {
      A[i] = lme(-Lam[i], -Lam[i + 1]);
      double b = (tb + inv) - (te + inv) * exp(-Lam[i + 1] + Lam[i]);
      B[i] = log(b) - Lam[i];
    } else { A[i] = neg_inf(); B[i] = neg_inf(); }
  } else {
    if (r > 0) { A[i] = -Lam[i]; B[i] = log(ep[i] + 1.0 / r) - Lam[i]; }
    else { A[i] = neg_inf(); B[i] = neg_inf(); }
  }
}

// get_tint with age_begin == age_end (coal_EM.cpp:60-95): grid points before the copies of t
__device__ __forceinline__ int tint_k(int E, const double* ep, double t)
{
  for (int e = 0; e < E; e++) if (t < ep[e]) return e;
  return E;
}

// EM_shared(t, t, ...), coal_EM.cpp:153-295.  emit(e, num_e, denom_e) is called for every
// epoch in order (so callers can reduce across a warp); returns the log normaliser.
template <class Emit>
__device__ double task_shared(const EmCtx& c, double t, int k, bool active, Emit emit)
{
  const int E = c.E, et = k - 1;
  double num_t = 0, den_t = 0, nc = 1.0;
  bool good = false;
  if (active) {
    const double r = c.rate[et];
    const double c0 = c.Lam[et];
    const double c1 = c0 + r * (t - c.ep[et]);
    if (r > 0) {
      const double inv = 1.0 / r, tb = c.ep[et];
      num_t = lme(-c0, -c1);
      den_t = log((tb + inv) / inv - (t + inv) / inv * exp(-c1 + c0)) + log(inv) - c0;
    } else { num_t = neg_inf(); den_t = neg_inf(); }
    for (int e = 0; e <= et; e++) {
      const double v = (e < et) ? c.A[e] : num_t;
      if (nc == 1.0) nc = v; else nc = lse(nc, v);
    }
    good = !bad(nc);
  }
  double integ = 1.0;
  const int lim = (E - 1 < et + 1) ? E - 1 : et + 1;
  for (int e = 0; e < E; e++) {
    double ne = 0.0, de = 0.0;
    if (good) {
      if (e < lim) {
        ne = exp(((e < et) ? c.A[e] : num_t) - nc);
        if (integ > 0.0) integ -= ne; else integ = 0.0;
        de = exp(((e < et) ? c.B[e] : den_t) - nc);
        de += -c.ep[e] * ne + (c.ep[e + 1] - c.ep[e]) * integ;
        if (de < 0.0) de = 0.0;
      } else if (e == E - 1 && et == E - 1) {
        ne = exp(num_t - nc);
        de = exp(den_t - nc);
        de -= c.ep[e] * ne;
        if (de < 0.0) de = 0.0;
      }
    }
    emit(e, ne, de);
  }
  return good ? nc : 0.0;
}

// EM_notshared(t, t, ...), coal_EM.cpp:297-468 (327-357, 435-466)
template <class Emit>
__device__ double task_notshared(const EmCtx& c, double t, int k, bool active, Emit emit)
{
  const int E = c.E, et = k - 1;
  double num_t = 0, den_t = 0, nc = 0;
  bool good = false;
  if (active) {
    const double r = c.rate[et], inv = 1.0 / r;
    const double c1 = c.Lam[et] + r * (t - c.ep[et]);
    const double c2 = c1 + r * (t - t);
    if (et != E - 1) {
      const double c3 = c2 + r * (c.ep[k] - t);
      if (r > 0) {
        num_t = lme(-c2, -c3);
        den_t = log((t + inv) - (c.ep[k] + inv) * exp(-c3 + c2)) - c2;
        nc = num_t;
      } else { num_t = neg_inf(); den_t = neg_inf(); nc = neg_inf(); }
      for (int e = et + 1; e < E; e++) nc = lse(nc, c.A[e]);
    } else {
      num_t = -c2;
      den_t = log(t + inv) - c2;
      nc = num_t;
    }
    good = !bad(nc);
  }
  double integ = 1.0;
  for (int e = 0; e < E; e++) {
    double ne = 0.0, de = 0.0;
    if (good) {
      if (e < et) {
        de = c.ep[e + 1] - c.ep[e];
      } else {
        const double ln = (e == et) ? num_t : c.A[e];
        const double ld = (e == et) ? den_t : c.B[e];
        ne = exp(ln - nc);
        if (e < E - 1) {
          if (integ > 0.0) integ -= ne; else integ = 0.0;
          de = exp(ld - nc);
          de += -c.ep[e] * ne + (c.ep[e + 1] - c.ep[e]) * integ;
        } else {
          de = exp(ld - nc);
          de -= c.ep[e] * ne;
        }
        if (de < 0.0) de = 0.0;
      }
    }
    emit(e, ne, de);
  }
  return good ? nc : 0.0;
}

__device__ __forceinline__ double warp_sum(double v)
{
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- stage iii: one CTA per bootstrap replicate, EM to convergence -------------------------
__global__ void __launch_bounds__(EM_THREADS)
k_em(int E, const double* __restrict__ epochs, const double* __restrict__ rates_init,
     const double* __restrict__ age_bin_g, const double* __restrict__ counts, int max_iter,
     double* __restrict__ rates_out, int32_t* __restrict__ iters_out, double* __restrict__ ll_out)
{
  extern __shared__ double sm[];
  double* ep = sm;                 // [E]
  double* rate = ep + E;           // [E]
  double* Lam = rate + E;          // [E]
  double* A = Lam + E;             // [E]
  double* B = A + E;               // [E]
  double* wn = B + E;              // [EM_WARPS][E]
  double* wd = wn + EM_WARPS * E;  // [EM_WARPS][E]
  double* wl = wd + EM_WARPS * E;  // [EM_WARPS]
  double* tn = wl + EM_WARPS;      // [E]
  double* td = tn + E;             // [E]
  __shared__ int stop_flag;
  __shared__ double ll_s, prev_s;

  const int rep = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_shared = tid < EM_HALF;
  const int bin = is_shared ? tid : tid - EM_HALF;
  for (int e = tid; e < E; e += blockDim.x) { ep[e] = epochs[e]; rate[e] = rates_init[e]; }
  if (tid == 0) { stop_flag = 0; ll_s = neg_inf(); }
  __syncthreads();
  double t = 0.0, cnt = 0.0;
  int k = 1;
  if (bin < NBINS) {
    t = age_bin_g[bin];
    k = tint_k(E, ep, t);
    cnt = counts[(size_t)rep * 2 * NBINS + (is_shared ? 0 : NBINS) + bin];
  }
  const bool active = bin < NBINS && cnt > 0;  // coal.cpp:3706, 3719
  EmCtx c{E, ep, rate, A, B, Lam};

  int iter = 0;
  for (; iter < max_iter; iter++) {
    if (tid == 0) em_cumhaz(E, ep, rate, Lam);
    __syncthreads();
    for (int e = tid; e < E; e += blockDim.x) em_AB(E, ep, rate, Lam, e, A, B);
    __syncthreads();
    auto emit = [&](int e, double ne, double de) {
      double a = warp_sum(cnt * ne), b = warp_sum(cnt * de);
      if (lane == 0) { wn[warp * E + e] = a; wd[warp * E + e] = b; }
    };
    double logl = is_shared ? task_shared(c, t, k, active, emit) : task_notshared(c, t, k, active, emit);
    double l = warp_sum(active ? cnt * logl : 0.0);
    if (lane == 0) wl[warp] = l;
    __syncthreads();
    for (int e = tid; e < E; e += blockDim.x) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < EM_WARPS; w++) { a += wn[w * E + e]; b += wd[w * E + e]; }
      tn[e] = a; td[e] = b;
    }
    __syncthreads();
    if (tid == 0) {
      double ll = 0.0;
      for (int w = 0; w < EM_WARPS; w++) ll += wl[w];
      for (int e = 0; e < E; e++) {  // M-step, coal.cpp:3771-3815 (regularise == 2)
        if (tn[e] == 0) rate[e] = (e > 0) ? rate[e - 1] : 0.0;
        else if (td[e] == 0) { }
        else { double r = tn[e] / td[e]; rate[e] = (r < 5e-9) ? 5e-9 : r; }
      }
      prev_s = ll_s;
      ll_s = ll;
      if ((ll / prev_s > 1.0 - 1e-7) && (iter > 1000)) stop_flag = 1;  // coal.cpp:3822
    }
    __syncthreads();
    if (stop_flag) break;
  }
  for (int e = tid; e < E; e += blockDim.x) rates_out[(size_t)rep * E + e] = rate[e];
  if (tid == 0) { iters_out[rep] = iter; ll_out[rep] = ll_s; }
}

// E-step probe: one thread per age, plain stores
__global__ void k_estep(int shared, int E, const double* __restrict__ epochs, const double* __restrict__ rates,
                        int n_t, const double* __restrict__ tt, double* __restrict__ num, double* __restrict__ denom,
                        double* __restrict__ logl)
{
  extern __shared__ double sm[];
  double* ep = sm; double* rate = ep + E; double* Lam = rate + E; double* A = Lam + E; double* B = A + E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) { ep[e] = epochs[e]; rate[e] = rates[e]; }
  __syncthreads();
  if (threadIdx.x == 0) em_cumhaz(E, ep, rate, Lam);
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) em_AB(E, ep, rate, Lam, e, A, B);
  __syncthreads();
  EmCtx c{E, ep, rate, A, B, Lam};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_t; i += gridDim.x * blockDim.x) {
    const double t = tt[i];
    const int k = tint_k(E, ep, t);
    double* n = num + (size_t)i * E;
    double* d = denom + (size_t)i * E;
    auto emit = [&](int e, double ne, double de) { n[e] = ne; d[e] = de; };
    logl[i] = shared ? task_shared(c, t, k, true, emit) : task_notshared(c, t, k, true, emit);
  }
}

// ---- launchers -------------------------------------------------------------------------------
int run_bootstrap(colate_handle* h, int R, int num_blocks, double age)
{
  k_bootstrap<<<R, 256, 0, h->stream>>>(num_blocks, h->d_weights.as<int32_t>(), h->d_blockstats.as<double>(), age,
                                        h->d_agebin.as<double>(), h->d_counts.as<double>());
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int run_em(colate_handle* h, int R, int E, int max_iter)
{
  const size_t smem = sizeof(double) * ((size_t)5 * E + 2 * EM_WARPS * E + EM_WARPS + 2 * E);
  CK(cudaFuncSetAttribute(k_em, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
  k_em<<<R, EM_THREADS, smem, h->stream>>>(E, h->d_epochs.as<double>(), h->d_rates.as<double>() /*init*/,
                                           h->d_agebin.as<double>(), h->d_counts.as<double>(), max_iter,
                                           h->d_rates.as<double>() + E, h->d_iters.as<int32_t>(), h->d_ll.as<double>());
  h->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int run_estep(colate_handle* h, int shared, int E, int n_t)
{
  const size_t smem = sizeof(double) * 5 * E;
  double* base = h->d_tmp.as<double>();  // [t: n_t][num: n_t*E][denom: n_t*E][logl: n_t]
  k_estep<<<(n_t + 127) / 128, 128, smem, h->stream>>>(shared, E, h->d_epochs.as<double>(), h->d_rates.as<double>(), n_t, base,
                                                       base + n_t, base + n_t + (size_t)n_t * E,
                                                       base + n_t + 2 * (size_t)n_t * E);
  CK(cudaGetLastError());
  return 0;
}

}  // namespace colate
