/* Fast writer of synthetic Relate .mut text for the benchmarks (colate_b200/synth.py: write_mut_fast).
 * Produces, byte for byte, what synth.write_mut() writes row by row in Python (format of
 * Mutations::Dump, include/src/mutations.cpp:298-328): ages as the shortest decimal that strtof reads
 * back to the same float, in positional notation.  Not part of the hot path and not linked into
 * libcolate_b200.so; built as colate_b200/libsynthio.so. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* shortest round-trip decimal of a float, positional, no trailing zeros / point */
static int fmt_f32(float x, char* out)
{
  if (x == 0.0f) { out[0] = '0'; return 1; }
  char e[40], t[40];
  /* the correctly rounded p-digit decimal reads back as x for every p from some p0 <= 9 on: bisect p0 */
  int lo = 1, hi = 9;
  snprintf(e, sizeof e, "%.8e", (double)x);
  while (lo < hi) {
    const int p = (lo + hi) / 2;
    snprintf(t, sizeof t, "%.*e", p - 1, (double)x);
    if (strtof(t, NULL) == x) { hi = p; memcpy(e, t, sizeof e); } else lo = p + 1;
  }
  /* e = [-]d[.ddd]e[+-]XX */
  char digits[16];
  int nd = 0, neg = 0;
  const char* s = e;
  if (*s == '-') { neg = 1; s++; }
  for (; *s && *s != 'e'; s++) if (*s >= '0' && *s <= '9') digits[nd++] = *s;
  int ex = atoi(s + 1);
  while (nd > 1 && digits[nd - 1] == '0') nd--;
  int n = 0;
  if (neg) out[n++] = '-';
  if (ex >= nd - 1) {
    memcpy(out + n, digits, nd); n += nd;
    for (int i = 0; i < ex - (nd - 1); i++) out[n++] = '0';
  } else if (ex >= 0) {
    memcpy(out + n, digits, ex + 1); n += ex + 1;
    out[n++] = '.';
    memcpy(out + n, digits + ex + 1, nd - ex - 1); n += nd - ex - 1;
  } else {
    out[n++] = '0'; out[n++] = '.';
    for (int i = 0; i < -ex - 1; i++) out[n++] = '0';
    memcpy(out + n, digits, nd); n += nd;
  }
  return n;
}

static int fmt_i64(long long v, char* out)
{
  char tmp[24];
  int n = 0, neg = v < 0;
  unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  int k = 0;
  if (neg) out[k++] = '-';
  while (n) out[k++] = tmp[--n];
  return k;
}

static const char HEADER[] =
    "snp;pos_of_snp;dist;rs-id;tree_index;branch_indices;is_not_mapping;is_flipped;age_begin;age_end;"
    "ancestral_allele/alternative_allele;upstream_allele;downstream_allele;\n";

/* rows [lo, hi) of the site arrays -> text in `out` (capacity cap); returns bytes written, -1 if cap is too small */
long long synth_mut_text(long long lo, long long hi, const int32_t* pos, const float* ab, const float* ae, const uint8_t* flipped,
                         const int32_t* n_branch, const uint8_t* anc, const uint8_t* der, const uint8_t* odd, char* out, long long cap)
{
  long long n = 0;
  const long long hl = (long long)sizeof HEADER - 1;
  if (cap < hl) return -1;
  memcpy(out, HEADER, hl);
  n = hl;
  for (long long i = lo; i < hi; i++) {
    if (cap - n < 200) return -1;
    const long long k = i - lo;
    char* o = out + n;
    int m = 0;
    m += fmt_i64(k, o + m); o[m++] = ';';
    m += fmt_i64(pos[i], o + m); o[m++] = ';';
    m += fmt_i64(i + 1 < hi ? (long long)pos[i + 1] - pos[i] : 1, o + m); o[m++] = ';';
    o[m++] = 'r'; o[m++] = 's'; m += fmt_i64(k, o + m); o[m++] = ';';
    m += fmt_i64(k / 3, o + m); o[m++] = ';';
    if (n_branch[i] == 1) { o[m++] = '1'; o[m++] = '7'; } else { memcpy(o + m, "17 23", 5); m += 5; }
    o[m++] = ';';
    o[m++] = n_branch[i] > 1 ? '1' : '0'; o[m++] = ';';
    m += fmt_i64(flipped[i], o + m); o[m++] = ';';
    m += fmt_f32(ab[i], o + m); o[m++] = ';';
    m += fmt_f32(ae[i], o + m); o[m++] = ';';
    if (odd[i] == 1) { o[m++] = (char)anc[i]; o[m++] = 'T'; o[m++] = '/'; o[m++] = (char)der[i]; }
    else if (odd[i] == 2) { o[m++] = 'N'; o[m++] = 'A'; }
    else { o[m++] = (char)anc[i]; o[m++] = '/'; o[m++] = (char)der[i]; }
    memcpy(o + m, ";A;C;\n", 6); m += 6;
    n += m;
  }
  return n;
}
