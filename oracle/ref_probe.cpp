// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Stage-level probe around the UNMODIFIED reference sources.  This TU pulls in
// /root/reference/include/coal/coal.cpp by #include (the reference itself builds
// Colate.cpp that way, Colate.cpp:1) and exposes a few extern "C" entry points so
// tests can call the reference's own parse_tmptmp / coal_EM / <random> directly.
// Built by oracle/Makefile into oracle/_ref/libcolate_ref.so (git-ignored).
// No reference source is copied into this repository.
#include "coal.cpp"
#include "coal_EM_old.hpp"

#include <sstream>
#include <cstring>

extern "C" {

// Calls parse_tmptmp (coal.cpp:2072) exactly as mut() does (coal.cpp:3317).
// out_* are [500*185] doubles; for the two emp matrices only row 0 (the first 185
// entries of the 185x185 matrix) is returned, emp_rest[0..1] get the sum of all
// other entries (expected 0).  mt_out receives the 624 state words + position as
// serialised by operator<<(std::mt19937).
int ref_parse_tmptmp(int n_chr, const char** chr_names, const char** mut_files,
                     const char* target_file, const char* ref_file,
                     const char** target_masks, const char** ref_masks,
                     double age, double ref_age, int seed,
                     double* out_shared, double* out_notshared,
                     double* out_shared_emp, double* out_notshared_emp,
                     double* emp_rest, unsigned int* mt_out /*[625]*/,
                     unsigned int* next_words /*[8]*/)
{
  double C = 10;
  int num_age_bins = ((int)(log(1e8) * C)) + 1;
  int num_bases_per_block = 30e6;
  int num_blocks = 500;
  std::vector<std::vector<double>> a(num_blocks), b(num_blocks), c(num_blocks), d(num_blocks);
  for (int i = 0; i < num_blocks; i++) {
    a[i].assign(num_age_bins, 0.0);
    b[i].assign(num_age_bins, 0.0);
    c[i].assign(num_age_bins * num_age_bins, 0.0);
    d[i].assign(num_age_bins * num_age_bins, 0.0);
  }
  std::vector<std::string> name_chr, filename_mut, tmask, rmask;
  for (int i = 0; i < n_chr; i++) {
    name_chr.push_back(chr_names[i]);
    filename_mut.push_back(mut_files[i]);
    if (target_masks) tmask.push_back(target_masks[i]);
    if (ref_masks) rmask.push_back(ref_masks[i]);
  }
  std::string ft(target_file), fr(ref_file);
  std::mt19937 rng;
  rng.seed(seed);
  int nb = parse_tmptmp(name_chr, filename_mut, ft, fr, tmask, rmask, age, ref_age, C, rng,
                        num_bases_per_block, a, b, c, d);
  emp_rest[0] = emp_rest[1] = 0.0;
  for (int i = 0; i < nb && i < 500; i++) {
    for (int k = 0; k < num_age_bins; k++) {
      out_shared[i * num_age_bins + k] = a[i][k];
      out_notshared[i * num_age_bins + k] = b[i][k];
      out_shared_emp[i * num_age_bins + k] = c[i][k];
      out_notshared_emp[i * num_age_bins + k] = d[i][k];
    }
    for (int k = num_age_bins; k < num_age_bins * num_age_bins; k++) {
      emp_rest[0] += c[i][k];
      emp_rest[1] += d[i][k];
    }
  }
  std::stringstream ss;
  ss << rng;
  for (int i = 0; i < 625; i++) {
    unsigned long v;
    ss >> v;
    mt_out[i] = (unsigned int)v;
  }
  for (int i = 0; i < 8; i++) next_words[i] = (unsigned int)rng();
  return nb;
}

// coal_EM (coal_EM.hpp:38) + EM_shared / EM_notshared (coal_EM.cpp:153, 297).
double ref_em_shared(int E, const double* epochs, const double* rates, double t0, double t1,
                     double* num, double* denom)
{
  std::vector<double> ep(epochs, epochs + E), r(rates, rates + E), n(E, 0.0), d(E, 0.0);
  coal_EM EM(ep, r);
  double ll = EM.EM_shared(t0, t1, n, d);
  for (int e = 0; e < E; e++) { num[e] = n[e]; denom[e] = d[e]; }
  return ll;
}

double ref_em_notshared(int E, const double* epochs, const double* rates, double t0, double t1,
                        double* num, double* denom)
{
  std::vector<double> ep(epochs, epochs + E), r(rates, rates + E), n(E, 0.0), d(E, 0.0);
  coal_EM EM(ep, r);
  double ll = EM.EM_notshared(t0, t1, n, d);
  for (int e = 0; e < E; e++) { num[e] = n[e]; denom[e] = d[e]; }
  return ll;
}

// The independent E-step the reference's unit test checks coal_EM against
// (coal_EM_old.hpp:157-195, test_aDNA.cpp:68-212).
double ref_em_simplified_shared(int E, const double* epochs, const double* rates, double t,
                                double* num, double* denom)
{
  std::vector<double> ep(epochs, epochs + E), r(rates, rates + E), n(E, 0.0), d(E, 0.0);
  coal_EM_simplified EM(ep, r);
  double ll = EM.EM_shared(t, n, d);
  for (int e = 0; e < E; e++) { num[e] = n[e]; denom[e] = d[e]; }
  return ll;
}

double ref_em_simplified_notshared(int E, const double* epochs, const double* rates, double t,
                                   double* num, double* denom)
{
  std::vector<double> ep(epochs, epochs + E), r(rates, rates + E), n(E, 0.0), d(E, 0.0);
  coal_EM_simplified EM(ep, r);
  double ll = EM.EM_notshared(t, n, d);
  for (int e = 0; e < E; e++) { num[e] = n[e]; denom[e] = d[e]; }
  return ll;
}

// libstdc++ <random> as the reference uses it (coal.cpp:2077, 2262; 3330, 3355).
void ref_uniform_real(int seed, long discard, int n, double* out)
{
  std::mt19937 rng;
  rng.seed(seed);
  rng.discard(discard);
  std::uniform_real_distribution<double> d(0, 1);
  for (int i = 0; i < n; i++) out[i] = d(rng);
}

void ref_uniform_int(int seed, long discard, int num_blocks, int n, int* out)
{
  std::mt19937 rng;
  rng.seed(seed);
  rng.discard(discard);
  std::uniform_int_distribution<int> d(0, num_blocks - 1);
  for (int i = 0; i < n; i++) out[i] = d(rng);
}

void ref_mt_words(int seed, long discard, int n, unsigned int* out)
{
  std::mt19937 rng;
  rng.seed(seed);
  rng.discard(discard);
  for (int i = 0; i < n; i++) out[i] = (unsigned int)rng();
}

// Which overload does `log(10*float)` resolve to inside coal.cpp's include set
// (coal.cpp:2253)?  Returns sizeof of the result type: 4 = float logf, 8 = double.
int ref_log_float_arg_size()
{
  float x = 3.0f;
  return (int)sizeof(log(10 * x));
}

// The bin index expressions exactly as spelled at coal.cpp:2253 / 2265.
int ref_bin_of_float_age(float age_end)
{
  double C = 10;
  return std::max(0, (int)std::round(log(10 * age_end) * C) + 1);
}
int ref_bin_of_double_age(double a)
{
  double C = 10;
  return std::max(0, (int)std::round(log(10 * a) * C) + 1);
}


// Mutations::Read(filename) (include/src/mutations.cpp:259-283 -> 56-257), the reader parse_tmptmp uses (coal.cpp:2115-2116),
// on one .mut file; returns the rows it produced (or -1 if it threw: std::stoi / std::stof throw on fields they cannot
// convert, which ends the reference run) and, per row, the fields the tmp/tmp path looks at (coal.cpp:2150-2176).
// mutation_type is returned as the first 15 characters + its length.
long ref_read_mut(const char* filename, long cap, int* pos, float* age_begin, float* age_end, int* flipped, int* n_branch,
                  char* type15 /*[cap][16]*/, int* type_len)
{
  Mutations m;
  try {
    m.Read(std::string(filename));
  } catch (const std::exception& e) {
    return -1;
  }
  long n = (long)m.info.size();
  for (long i = 0; i < n && i < cap; i++) {
    const SNPInfo& s = m.info[i];
    pos[i] = s.pos;
    age_begin[i] = s.age_begin;
    age_end[i] = s.age_end;
    flipped[i] = s.flipped ? 1 : 0;
    n_branch[i] = (int)s.branch.size();
    memset(type15 + 16 * i, 0, 16);
    strncpy(type15 + 16 * i, s.mutation_type.c_str(), 15);
    type_len[i] = (int)s.mutation_type.size();
  }
  return n;
}

}  // extern "C"

extern "C" {

// parse_onebambam (coal.cpp:1799-2069) exactly as mut() calls it (coal.cpp:3288): the bam/bam front-end whose weighting
// variant is row N3 of SURVEY.md 8(f).  The two "BAM" files are oracle/hts_stubs.c fake BAMs (synthetic reads); everything
// above the htslib calls -- bam_parser's pileup ring, the filters, the weights -- is the reference's own code.
// Outputs as ref_parse_tmptmp.
int ref_parse_onebambam(const char* params, int n_chr, const char** chr_names, const char** mut_files,
                        const char* target_bam, const char* ref_bam, const char** target_masks, const char** ref_masks,
                        const char** ref_genomes, int seed,
                        double* out_shared, double* out_notshared, double* out_shared_emp, double* out_notshared_emp,
                        double* emp_rest, unsigned int* mt_out /*[625]*/)
{
  double C = 10;
  int num_age_bins = ((int)(log(1e8) * C)) + 1;
  int num_bases_per_block = 30e6;
  int num_blocks = 500;
  std::vector<std::vector<double>> a(num_blocks), b(num_blocks), c(num_blocks), d(num_blocks);
  for (int i = 0; i < num_blocks; i++) {
    a[i].assign(num_age_bins, 0.0);
    b[i].assign(num_age_bins, 0.0);
    c[i].assign(num_age_bins * num_age_bins, 0.0);
    d[i].assign(num_age_bins * num_age_bins, 0.0);
  }
  std::vector<std::string> name_chr, filename_mut, tmask, rmask, refg, ft(1, target_bam), fr(1, ref_bam);
  for (int i = 0; i < n_chr; i++) {
    name_chr.push_back(chr_names[i]);
    filename_mut.push_back(mut_files[i]);
    refg.push_back(ref_genomes[i]);
    if (target_masks) tmask.push_back(target_masks[i]);
    if (ref_masks) rmask.push_back(ref_masks[i]);
  }
  std::string prm(params);
  std::mt19937 rng;
  rng.seed(seed);
  int nb = parse_onebambam(prm, name_chr, filename_mut, ft, fr, tmask, rmask, refg, 0.0, 0.0, C, rng, num_bases_per_block, a, b, c, d);
  emp_rest[0] = emp_rest[1] = 0.0;
  for (int i = 0; i < nb && i < 500; i++) {
    for (int k = 0; k < num_age_bins; k++) {
      out_shared[i * num_age_bins + k] = a[i][k];
      out_notshared[i * num_age_bins + k] = b[i][k];
      out_shared_emp[i * num_age_bins + k] = c[i][k];
      out_notshared_emp[i * num_age_bins + k] = d[i][k];
    }
    for (int k = num_age_bins; k < num_age_bins * num_age_bins; k++) {
      emp_rest[0] += c[i][k];
      emp_rest[1] += d[i][k];
    }
  }
  std::stringstream ss;
  ss << rng;
  for (int i = 0; i < 625; i++) {
    unsigned long v;
    ss >> v;
    mt_out[i] = (unsigned int)v;
  }
  return nb;
}

// parse_vcfvcf (coal.cpp:907-1228) as mut() calls it (coal.cpp:3200): per-chromosome target / reference "bcf" files (fake BCFs of
// oracle/hts_stubs.c), optional masks and reference genome.  Outputs as ref_parse_tmptmp.
int ref_parse_vcfvcf(int n_chr, const char** mut_files, const char** target_bcf, const char** ref_bcf, const char** target_masks,
                     const char** ref_masks, const char** ref_genomes, int seed,
                     double* out_shared, double* out_notshared, double* out_shared_emp, double* out_notshared_emp,
                     double* emp_rest, unsigned int* mt_out /*[625]*/)
{
  double C = 10;
  int num_age_bins = ((int)(log(1e8) * C)) + 1;
  int num_bases_per_block = 30e6;
  int num_blocks = 500;
  std::vector<std::vector<double>> a(num_blocks), b(num_blocks), c(num_blocks), d(num_blocks);
  for (int i = 0; i < num_blocks; i++) {
    a[i].assign(num_age_bins, 0.0);
    b[i].assign(num_age_bins, 0.0);
    c[i].assign(num_age_bins * num_age_bins, 0.0);
    d[i].assign(num_age_bins * num_age_bins, 0.0);
  }
  std::vector<std::string> filename_mut, ft, fr, tmask, rmask, refg;
  for (int i = 0; i < n_chr; i++) {
    filename_mut.push_back(mut_files[i]);
    ft.push_back(target_bcf[i]);
    fr.push_back(ref_bcf[i]);
    if (target_masks) tmask.push_back(target_masks[i]);
    if (ref_masks) rmask.push_back(ref_masks[i]);
    if (ref_genomes) refg.push_back(ref_genomes[i]);
  }
  std::mt19937 rng;
  rng.seed(seed);
  int nb = parse_vcfvcf(filename_mut, ft, fr, tmask, rmask, refg, 0.0, 0.0, C, rng, num_bases_per_block, a, b, c, d);
  emp_rest[0] = emp_rest[1] = 0.0;
  for (int i = 0; i < nb && i < 500; i++) {
    for (int k = 0; k < num_age_bins; k++) {
      out_shared[i * num_age_bins + k] = a[i][k];
      out_notshared[i * num_age_bins + k] = b[i][k];
      out_shared_emp[i * num_age_bins + k] = c[i][k];
      out_notshared_emp[i * num_age_bins + k] = d[i][k];
    }
    for (int k = num_age_bins; k < num_age_bins * num_age_bins; k++) {
      emp_rest[0] += c[i][k];
      emp_rest[1] += d[i][k];
    }
  }
  std::stringstream ss;
  ss << rng;
  for (int i = 0; i < 625; i++) {
    unsigned long v;
    ss >> v;
    mt_out[i] = (unsigned int)v;
  }
  return nb;
}

// The pileup bam_parser holds at given positions: for every 1-based position bp[i] of chromosome `contig` (ascending),
// read_to_pos(bp - 1) as coal.cpp:1885-1888 does, then counts[i][0..3] = count_alleles[(bp-1) % num_entries] if that ring
// entry belongs to bp - 1, else zeros (and covered[i] = 0).  This is the "pre-decoded array" the N3 entry point takes.
int ref_bam_pileup(const char* params, const char* bam, const char* contig, const char* ref_genome, long n, const int* bp,
                   int* counts /*[n][4]*/, unsigned char* covered)
{
  std::string prm(params), fb(bam);
  bam_parser P(fb, prm);
  P.assign_contig(contig, ref_genome);
  for (long i = 0; i < n; i++) {
    const int q = bp[i] - 1;
    P.read_to_pos(q);
    const bool hit = P.pos_of_entry[q % P.num_entries] == q;
    covered[i] = hit ? 1 : 0;
    for (int k = 0; k < 4; k++) counts[4 * i + k] = hit ? P.count_alleles[q % P.num_entries][k] : 0;
  }
  return 0;
}

}  // extern "C"
