// Host side of the RNG contract: std::mt19937 as a "state window"
// generator, libstdc++'s uniform_int_distribution for the bootstrap draws
// (coal.cpp:3330, 3355), and the GF(2) machinery for jump-ahead: the characteristic
// polynomial of MT19937 (Berlekamp-Massey on its own output) and t^(200*2^q) mod p, the
// seed-independent polynomials the device jump kernel applies to reach every chunk of the
// reference's generator stream in parallel.
#include "internal.h"

#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace colate {

// ---- MT19937 on a window: w[0..623] are the 624 words preceding the next output ----------
static inline uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

uint32_t mt_temper(uint32_t z)
{
  z ^= (z >> 11);
  z ^= (z << 7) & 0x9d2c5680u;
  z ^= (z << 15) & 0xefc60000u;
  z ^= (z >> 18);
  return z;
}

uint32_t mt_untemper(uint32_t y)
{
  y ^= (y >> 18);
  y ^= (y << 15) & 0xefc60000u;
  // invert y ^= (y << 7) & 0x9d2c5680
  uint32_t t = y;
  for (int i = 0; i < 5; i++) t = y ^ ((t << 7) & 0x9d2c5680u);
  y = t;
  // invert y ^= y >> 11
  t = y;
  for (int i = 0; i < 3; i++) t = y ^ (t >> 11);
  return t;
}

void mt_seed_window(uint32_t seed, uint32_t* w)
{
  w[0] = seed;
  for (uint32_t i = 1; i < 624; i++) w[i] = 1812433253u * (w[i - 1] ^ (w[i - 1] >> 30)) + i;
}

// advance the window by 624 words in place (one full twist)
static void mt_twist_window(uint32_t* w)
{
  for (int k = 0; k < 624; k++) w[k] = mt_mix(w[k], w[(k + 1) % 624], w[(k + 397) % 624]);
}

void mt_generate_window(uint32_t* w, int64_t n, uint32_t* out)
{
  // slide the window one word at a time through a 1248-word scratch so that n need not be
  // a multiple of 624
  int64_t done = 0;
  while (done < n) {
    uint32_t nw[624];
    memcpy(nw, w, sizeof nw);
    mt_twist_window(nw);  // nw[k] = x[base+624+k]
    int64_t take = n - done < 624 ? n - done : 624;
    for (int64_t k = 0; k < take; k++) out[done + k] = mt_temper(nw[k]);
    if (take == 624) {
      memcpy(w, nw, sizeof nw);
    } else {
      uint32_t tmp[624];
      for (int k = 0; k < 624; k++) tmp[k] = (k + take < 624) ? w[k + take] : nw[k + take - 624];
      memcpy(w, tmp, sizeof tmp);
    }
    done += take;
  }
}

struct WindowGen {
  uint32_t* w;
  uint32_t buf[624];
  int pos = 624;
  explicit WindowGen(uint32_t* win) : w(win) {}
  uint32_t next()
  {
    if (pos == 624) { mt_generate_window(w, 624, buf); pos = 0; }
    return buf[pos++];
  }
  // put the unread part of buf back: the caller's window must end right after the last
  // word handed out
  void finish(int64_t consumed_total, const uint32_t* w_before)
  {
    memcpy(w, w_before, 624 * sizeof(uint32_t));
    std::vector<uint32_t> scratch((size_t)consumed_total);
    if (consumed_total > 0) mt_generate_window(w, consumed_total, scratch.data());
  }
};

// std::uniform_int_distribution<int>(0, n-1) for a 32-bit engine (libstdc++ 13, Lemire)
static int uniform_int(WindowGen& g, int n, int64_t& consumed)
{
  uint32_t range = (uint32_t)n;
  uint64_t prod = (uint64_t)g.next() * range;
  consumed++;
  uint32_t low = (uint32_t)prod;
  if (low < range) {
    uint32_t th = (uint32_t)(0u - range) % range;
    while (low < th) {
      prod = (uint64_t)g.next() * range;
      consumed++;
      low = (uint32_t)prod;
    }
  }
  return (int)(prod >> 32);
}

void draw_block_weights(uint32_t* w, int R, int num_blocks, int32_t* weights)
{
  if (R == 1) {  // coal.cpp:3350-3351: no draw at all
    for (int j = 0; j < num_blocks; j++) weights[j] = 1;
    return;
  }
  uint32_t before[624];
  memcpy(before, w, sizeof before);
  WindowGen g(w);
  int64_t consumed = 0;
  for (int i = 0; i < R; i++) {
    int32_t* row = weights + (size_t)i * num_blocks;
    for (int j = 0; j < num_blocks; j++) row[j] = 0;
    for (int j = 0; j < num_blocks; j++) row[uniform_int(g, num_blocks, consumed)] += 1;
  }
  g.finish(consumed, before);
}

// One used row with age_begin > 0 whose age interval reaches past the age grid, exactly as the reference's loop
// at coal.cpp:2279-2294 walks it: draw (std::uniform_real_distribution<double>(0,1) = generate_canonical: two
// engine words, (x1 + x2 * 2^32) / 2^64 with one rounding, clamped below 1), sampled_age = U * len + age_begin, bin
// = max(0,(int)round(log(10 * age) * 10) + 1); a draw whose bin reaches 185 is redrawn and does not count.
// cnt[slot] = accepted samples per bin (slot = bin), 100 in total.  Advances the state window by the words consumed
// and returns the number of redraws; -1: more than `max_redraws` (the interval barely touches the grid).
int64_t sample_deep_row_host(uint32_t* w, double age_begin, double len, uint8_t* cnt /*[192]*/, int64_t max_redraws)
{
  uint32_t before[624];
  memcpy(before, w, sizeof before);
  WindowGen g(w);
  int64_t consumed = 0, redraws = 0;
  memset(cnt, 0, 192);
  int j = 0;
  while (j < COLATE_NUM_SAMPLES) {
    const uint32_t x1 = g.next(), x2 = g.next();
    consumed += 2;
    double u = ((double)x1 + (double)x2 * 4294967296.0) / 18446744073709551616.0;
    if (u >= 1.0) u = 0x1.fffffffffffffp-1;
    const double a = u * len + age_begin;
    const int b = bin_of_x10_host(10.0 * a);
    if (b >= NBINS) { if (++redraws > max_redraws) return -1; continue; }
    cnt[b]++;
    j++;
  }
  g.finish(consumed, before);
  return redraws;
}

// ---- GF(2)[t] ------------------------------------------------------------------------
static const int DEG = 19937;
static const int NW64 = 312;  // 19968 bits

struct Gf2 {
  std::vector<int> p_terms;             // exponents of p(t) below t^19937
  std::vector<std::vector<uint32_t>> g; // g[q] = t^(200*2^q) mod p, 624 x u32, bit i = coeff of t^i
  bool ok = false;
};
static Gf2 G;
static std::mutex G_mu;

static inline int getbit(const uint64_t* a, int i) { return (a[i >> 6] >> (i & 63)) & 1; }
static inline void flipbit(uint64_t* a, int i) { a[i >> 6] ^= (1ull << (i & 63)); }

// Berlekamp-Massey on bit 0 of x[1..], the untempered MT19937 sequence of seed 5489
static bool compute_charpoly(std::vector<int>& terms)
{
  const int N = 2 * DEG + 64;
  uint32_t w[624];
  mt_seed_window(5489u, w);
  std::vector<uint8_t> s(N);
  {
    std::vector<uint32_t> out(N + 1);
    // untempered sequence: regenerate through the window and undo the tempering
    mt_generate_window(w, N + 1, out.data());
    for (int n = 0; n < N; n++) s[n] = mt_untemper(out[n + 1]) & 1u;
  }
  const int W = (DEG + 64) / 64 + 2;
  std::vector<uint64_t> C(W, 0), B(W, 0), T(W), win(W, 0);
  C[0] = 1; B[0] = 1;
  int L = 0, m = 1;
  for (int n = 0; n < N; n++) {
    // win bit i = s[n-i]
    for (int k = W - 1; k > 0; k--) win[k] = (win[k] << 1) | (win[k - 1] >> 63);
    win[0] = (win[0] << 1) | s[n];
    uint64_t acc = 0;
    int lw = L / 64 + 1;
    for (int k = 0; k < lw; k++) acc ^= C[k] & win[k];
    int d = __builtin_parityll(acc);
    if (d == 0) { m++; continue; }
    bool grow = 2 * L <= n;
    if (grow) T = C;
    int ws = m >> 6, bs = m & 63;
    for (int k = W - 1; k >= ws; k--) {
      uint64_t v = B[k - ws] << bs;
      if (bs && k - ws - 1 >= 0) v |= B[k - ws - 1] >> (64 - bs);
      C[k] ^= v;
    }
    if (grow) { L = n + 1 - L; B = T; m = 1; } else m++;
  }
  if (L != DEG) return false;
  // p_k = c_{L-k}
  terms.clear();
  for (int k = 0; k < DEG; k++) if (getbit(C.data(), DEG - k)) terms.push_back(k);
  return getbit(C.data(), 0) == 1;
}

static void reduce_mod_p(std::vector<uint64_t>& a /* 2*NW64 words */, const std::vector<int>& terms)
{
  for (int bit = 2 * DEG - 2; bit >= DEG; bit--) {
    if (!getbit(a.data(), bit)) continue;
    flipbit(a.data(), bit);
    int sh = bit - DEG;
    for (int e : terms) flipbit(a.data(), sh + e);
  }
}

static void square_mod_p(const std::vector<uint64_t>& in, std::vector<uint64_t>& out, const std::vector<int>& terms)
{
  std::vector<uint64_t> a(2 * NW64 + 2, 0);
  for (int k = 0; k < NW64; k++) {
    uint64_t v = in[k];
    // spread the 64 bits of v over 128 bits
    uint64_t lo = v & 0xffffffffull, hi = v >> 32;
    auto spread = [](uint64_t x) {
      x = (x | (x << 16)) & 0x0000ffff0000ffffull;
      x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
      x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
      x = (x | (x << 2)) & 0x3333333333333333ull;
      x = (x | (x << 1)) & 0x5555555555555555ull;
      return x;
    };
    a[2 * k] = spread(lo);
    a[2 * k + 1] = spread(hi);
  }
  reduce_mod_p(a, terms);
  out.assign(a.begin(), a.begin() + NW64);
}

static bool gf2_init_locked(int q_max)
{
  if (!G.ok) {
    if (!compute_charpoly(G.p_terms)) return false;
    G.ok = true;
  }
  while ((int)G.g.size() <= q_max) {
    std::vector<uint64_t> cur(NW64, 0), nxt;
    if (G.g.empty()) {
      flipbit(cur.data(), 200);  // t^200, already reduced (200 < 19937)
      nxt = cur;
    } else {
      const std::vector<uint32_t>& prev = G.g.back();
      for (int k = 0; k < NW64; k++) cur[k] = (uint64_t)prev[2 * k] | ((uint64_t)prev[2 * k + 1] << 32);
      square_mod_p(cur, nxt, G.p_terms);
    }
    std::vector<uint32_t> g32(624);
    for (int k = 0; k < NW64; k++) { g32[2 * k] = (uint32_t)nxt[k]; g32[2 * k + 1] = (uint32_t)(nxt[k] >> 32); }
    G.g.push_back(std::move(g32));
  }
  return true;
}

// t^(200*2^q) mod p as 624 u32 words; nullptr on failure.  Pointer stays valid (vector of
// vectors only grows at the back and the inner buffers never move).
const uint32_t* jump_poly(int q)
{
  std::lock_guard<std::mutex> lk(G_mu);
  if (q < 0 || q > 48) return nullptr;
  if (!gf2_init_locked(q)) return nullptr;
  return G.g[q].data();
}

// t^(3 * 200 * 2^q) mod p = g[q] * g[q + 1] mod p: the third polynomial of a radix-4 level of the jump tree
// (kernels_mt.cu).  Schoolbook product over GF(2), then the reduction; cached.
const uint32_t* jump_poly3(int q)
{
  std::lock_guard<std::mutex> lk(G_mu);
  if (q < 0 || q > 47) return nullptr;
  if (!gf2_init_locked(q + 1)) return nullptr;
  static std::map<int, std::vector<uint32_t>> cache;
  auto it = cache.find(q);
  if (it != cache.end()) return it->second.data();
  std::vector<uint64_t> a(NW64, 0), b(NW64, 0), prod(2 * NW64 + 2, 0);
  for (int k = 0; k < NW64; k++) {
    a[k] = (uint64_t)G.g[q][2 * k] | ((uint64_t)G.g[q][2 * k + 1] << 32);
    b[k] = (uint64_t)G.g[q + 1][2 * k] | ((uint64_t)G.g[q + 1][2 * k + 1] << 32);
  }
  for (int i = 0; i < DEG; i++) {
    if (!getbit(a.data(), i)) continue;
    const int ws = i >> 6, bs = i & 63;
    for (int k = 0; k < NW64; k++) {
      prod[k + ws] ^= b[k] << bs;
      if (bs) prod[k + ws + 1] ^= b[k] >> (64 - bs);
    }
  }
  reduce_mod_p(prod, G.p_terms);
  std::vector<uint32_t> g32(624);
  for (int k = 0; k < NW64; k++) { g32[2 * k] = (uint32_t)prod[k]; g32[2 * k + 1] = (uint32_t)(prod[k] >> 32); }
  return cache.emplace(q, std::move(g32)).first->second.data();
}

int charpoly_terms(int* out, int cap)
{
  std::lock_guard<std::mutex> lk(G_mu);
  if (!gf2_init_locked(0)) return -1;
  int n = (int)G.p_terms.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = G.p_terms[i];
  return n;
}

// host reference of the jump (used by tests and for small inputs' sanity checks):
// window at offset +200*2^q words from `w`
void jump_window_host(const uint32_t* w, int q, uint32_t* out)
{
  const uint32_t* g = jump_poly(q);
  std::vector<uint32_t> X(DEG + 624 + 8);
  memcpy(X.data(), w, 624 * sizeof(uint32_t));
  for (size_t n = 624; n < X.size(); n++) X[n] = mt_mix(X[n - 624], X[n - 623], X[n - 227]);
  for (int j = 0; j < 624; j++) out[j] = 0;
  for (int i = 0; i < DEG; i++) {
    if (!((g[i >> 5] >> (i & 31)) & 1u)) continue;
    for (int j = 0; j < 624; j++) out[j] ^= X[i + j];
  }
}

}  // namespace colate

extern "C" {

void colate_mt_seed(uint32_t seed, uint32_t* mt_state) { colate::mt_seed_window(seed, mt_state); }

void colate_mt_generate(uint32_t* mt_state, int64_t n, uint32_t* out) { colate::mt_generate_window(mt_state, n, out); }

void colate_draw_block_weights(uint32_t* mt_state, int R, int num_blocks, int32_t* block_weights)
{
  colate::draw_block_weights(mt_state, R, num_blocks, block_weights);
}

// test hooks (declared in internal.h only)
int colate_test_charpoly_terms(int* out, int cap) { return colate::charpoly_terms(out, cap); }
int colate_test_jump_window_host(const uint32_t* w, int q, uint32_t* out)
{
  if (!colate::jump_poly(q)) return -1;
  colate::jump_window_host(w, q, out);
  return 0;
}
}
