// `Colate` command-line host, `--mode mut` on two precomputed .colate.in files.
// Mirrors the reference CLI (include/coal/Colate.cpp:6-116, flag table 11-45) and the control
// flow of mut() (include/coal/coal.cpp:3071-3863); all computation goes through the C ABI
// (include/colate_b200.h) onto the GPU.  There is no CPU fallback.
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/colate_b200.h"

namespace {

struct Options {
  std::map<std::string, std::string> kv;
  int count(const std::string& k) const { return kv.count(k) ? 1 : 0; }
  const std::string& get(const std::string& k) const { return kv.at(k); }
};

// option names of the reference's table (Colate.cpp:11-45); value-less: help, strandfilter
const char* kValueOpts[] = {"mode", "anc", "mut", "target_bcf", "reference_bcf", "target_mask", "reference_mask", "target_table",
                            "target_bam", "reference_bam", "target_tmp", "reference_tmp", "target_age", "reference_age", "ref_genome",
                            "anc_genome", "mask", "mask_cutoff", "chr", "bins", "lineage_bin", "outgroup_tmrca", "years_per_gen",
                            "coal", "seed", "num_bootstraps", "filters", "groups", "poplabels", "map", "input", "output",
                            // additions of this build
                            "num_bootstrap", "device"};

void help()
{
  std::cout << "Usage:\n  Colate --mode mut --mut <prefix> --target_tmp <t.colate.in> --reference_tmp <r.colate.in> --bins x,y,step\n"
               "         [--chr <file>] [--target_mask <prefix>] [--reference_mask <prefix>] [--target_age <years>]\n"
               "         [--reference_age <years>] [--years_per_gen <float>] [--coal <file>] [--seed <int>]\n"
               "         [--num_bootstraps <int>] [--device <int>] [--host_parse] -o <output prefix>\n"
            << std::endl;
}

bool parse(int argc, char** argv, Options& o)
{
  std::set<std::string> val(std::begin(kValueOpts), std::end(kValueOpts));
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    std::string key, value;
    bool has_value = false;
    if (a.rfind("--", 0) == 0) {
      key = a.substr(2);
      size_t eq = key.find('=');
      if (eq != std::string::npos) { value = key.substr(eq + 1); key = key.substr(0, eq); has_value = true; }
    } else if (a == "-o") key = "output";
    else if (a == "-i") key = "input";
    else { std::cerr << "Option '" << a << "' does not exist" << std::endl; return false; }
    if (key == "help" || key == "strandfilter" || key == "host_parse") { o.kv[key] = "1"; continue; }   // host_parse: addition of this build
    if (!val.count(key)) {  // cxxopts::option_not_exists_exception in the reference (cxxopts.hpp:1180-1185)
      std::cerr << "Option '" << key << "' does not exist" << std::endl;
      return false;
    }
    if (!has_value) {
      if (i + 1 >= argc) { std::cerr << "Option '" << key << "' is missing an argument" << std::endl; return false; }
      value = argv[++i];
    }
    o.kv[key] = value;
  }
  return true;
}

int die(const std::string& what)
{
  std::cerr << what << ": " << colate_last_error() << std::endl;
  return 1;
}

bool file_exists(const std::string& p) { std::ifstream f(p); return f.good(); }

// COLATE_TIMING=1: wall-clock phases on stderr
struct Phases {
  bool on = getenv("COLATE_TIMING") != nullptr;
  double t0 = now(), last = t0;
  static double now() { timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }
  void tick(const char* what) { if (!on) return; const double t = now(); fprintf(stderr, "[timing] %-28s %8.3f s (total %.3f)\n", what, t - last, t - t0); last = t; }
};

int run_mut(const Options& options)
{
  if (!options.count("mut") || !options.count("output")) {  // coal.cpp:3077-3086
    std::cout << "Not enough arguments supplied." << std::endl;
    std::cout << "Needed: mut, bins, output. Optional: target_tmp, reference_tmp, target_bcf, reference_bcf, target_bam, "
                 "reference_bam, ref_genome, target_age, reference_age, target_mask, reference_mask, coal, num_bootstrap, filters."
              << std::endl;
    help();
    std::cout << "Calculate coalescence rates for sample." << std::endl;
    exit(0);
  }
  std::cerr << "---------------------------------------------------------" << std::endl;
  std::cerr << "Calculating coalescence rates for (ancient) samples.." << std::endl;

  double ypg = 28.0;
  const bool has_ypg = options.count("years_per_gen");
  double age = colate_age_generations(options.count("target_age") ? options.get("target_age").c_str() : nullptr,
                                      options.count("reference_age") ? options.get("reference_age").c_str() : nullptr, has_ypg,
                                      has_ypg ? strtof(options.get("years_per_gen").c_str(), nullptr) : 0.0f, &ypg);
  std::cerr << age << std::endl;  // coal.cpp:3119
  const bool is_ancient = age > 0.0;
  std::cerr << "num_bins: " << COLATE_NUM_AGE_BINS << std::endl;

  int seed = (int)(std::time(0) + getpid());  // coal.cpp:3158
  if (options.count("seed")) seed = atoi(options.get("seed").c_str());
  int R = 1;  // coal.cpp:3164; the README spells the flag --num_bootstrap, the code --num_bootstraps
  if (options.count("num_bootstraps")) R = atoi(options.get("num_bootstraps").c_str());
  else if (options.count("num_bootstrap")) R = atoi(options.get("num_bootstrap").c_str());
  if (R < 0) R = 0;
  const std::string out = options.get("output");

  Phases ph;
  // the CUDA context comes up (about a second) while the main thread reads the input files
  colate_handle* h = nullptr;
  int create_rc = 0;
  std::string create_err;
  const int device = options.count("device") ? atoi(options.get("device").c_str()) : 0;
  std::thread init([&] { create_rc = colate_create(device, &h); if (create_rc) create_err = colate_last_error(); });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{init};
  auto wait_for_device = [&]() -> bool {
    if (init.joinable()) init.join();
    if (create_rc) std::cerr << "colate_create: " << create_err << std::endl;
    return create_rc == 0;
  };
  std::vector<double> counts((size_t)std::max(R, 1) * 2 * COLATE_NUM_AGE_BINS, 0.0);
  uint32_t mt[COLATE_MT_WORDS];
  colate_mt_seed((uint32_t)seed, mt);

  if (file_exists(out + ".colate_mat")) {  // stale-cache behaviour of coal.cpp:3169-3170, 3471-3499
    std::cerr << "Loading precomputed file " << out << ".colate_mat" << std::endl;
    std::ifstream is(out + ".colate_mat");
    double dummy;
    for (int b = 0; b < COLATE_NUM_AGE_BINS; b++) is >> dummy;
    for (int i = 0; i < R; i++)
      for (int k = 0; k < 2 * COLATE_NUM_AGE_BINS; k++) is >> counts[(size_t)i * 2 * COLATE_NUM_AGE_BINS + k];
  } else {
    if (!(options.count("target_tmp") && options.count("reference_tmp"))) {
      std::cerr << "This build reads --target_tmp/--reference_tmp (.colate.in) inputs only; bcf/bam front-ends are out of scope." << std::endl;
      return 1;
    }
    // file lists, coal.cpp:3290-3316
    std::vector<std::string> name_chr, f_mut, f_tmask, f_rmask;
    if (options.count("chr")) {
      std::ifstream is(options.get("chr"));
      if (is.fail()) std::cerr << "Error while opening file " << options.get("chr") << std::endl;
      std::string line;
      while (std::getline(is, line)) {
        name_chr.push_back(line);
        f_mut.push_back(options.get("mut") + "_chr" + line + ".mut");
        if (options.count("target_mask")) f_tmask.push_back(options.get("target_mask") + "_chr" + line + ".fa");
        if (options.count("reference_mask")) f_rmask.push_back(options.get("reference_mask") + "_chr" + line + ".fa");
      }
    } else {
      name_chr.push_back("");
      f_mut.push_back(options.get("mut"));
      if (options.count("target_mask")) f_tmask.push_back(options.get("target_mask"));
      if (options.count("reference_mask")) f_rmask.push_back(options.get("reference_mask"));
    }
    const int n_chr = (int)name_chr.size();
    {
      std::set<std::string> uniq(name_chr.begin(), name_chr.end());
      if ((int)uniq.size() != n_chr) { std::cerr << "Duplicate chromosome names in --chr" << std::endl; return 1; }
    }
    // readers -> SoA.  Plain-text .mut files are parsed on the GPU (colate_ingest_*: the bytes go to the device,
    // one thread per row); if any file only exists as .gz the host reader (zlib) takes over for all of them.
    std::vector<int64_t> site_off(n_chr + 1, 0);
    std::vector<int32_t> pos;
    std::vector<std::vector<char>> texts(n_chr);
    bool all_plain = !options.count("host_parse");
    for (int c = 0; c < n_chr && all_plain; c++) {
      FILE* f = fopen(f_mut[c].c_str(), "rb");
      if (!f) { all_plain = false; break; }
      fseek(f, 0, SEEK_END);
      const long sz = ftell(f);
      fseek(f, 0, SEEK_SET);
      texts[c].resize((size_t)std::max(sz, 0L));
      if (sz > 0 && fread(texts[c].data(), 1, (size_t)sz, f) != (size_t)sz) all_plain = false;
      fclose(f);
    }
    // the two .colate.in files: decoded on the host in one pass each (at least 19 bytes per record)
    std::vector<const char*> names;
    for (auto& s : name_chr) names.push_back(s.c_str());
    const std::string files[2] = {options.get("target_tmp"), options.get("reference_tmp")};
    struct GenomeHost { int64_t n = 0; std::vector<int32_t> rc, bp, aaf, daf; std::vector<uint16_t> al; } gh[2];
    for (int g = 0; g < 2; g++) {
      struct stat st;
      const int64_t cap = stat(files[g].c_str(), &st) == 0 ? (int64_t)st.st_size / 19 + 16 : 16;
      gh[g].rc.resize(cap); gh[g].bp.resize(cap); gh[g].aaf.resize(cap); gh[g].daf.resize(cap); gh[g].al.resize(cap);
      int64_t n = colate_read_colate_in(files[g].c_str(), n_chr, names.data(), cap, gh[g].rc.data(), gh[g].bp.data(), gh[g].aaf.data(),
                                        gh[g].daf.data(), gh[g].al.data());
      if (n < 0) { std::cerr << colate_last_error() << std::endl; n = 0; }  // the reference only warns (coal.cpp:2093-2098)
      gh[g].n = n;
    }
    ph.tick("read input files");
    if (!wait_for_device()) return 1;
    ph.tick("wait for the CUDA context");
    if (all_plain) {
      int64_t cap = 0;
      for (auto& t : texts) cap += (int64_t)std::count(t.begin(), t.end(), '\n') + 1;
      if (colate_ingest_begin(h, n_chr, cap)) return die("colate_ingest_begin");
      for (int c = 0; c < n_chr; c++) {
        std::cerr << "parsing CHR: " << c + 1 << " / " << n_chr << std::endl;
        const int64_t n = colate_ingest_mut_text(h, texts[c].data(), (int64_t)texts[c].size(), 0);
        if (n < 0) { std::cerr << colate_last_error() << std::endl; exit(1); }
        site_off[c + 1] = site_off[c] + n;
        std::vector<char>().swap(texts[c]);
      }
      if (colate_ingest_end(h)) return die("colate_ingest_end");
      if (options.count("target_mask") || options.count("reference_mask")) {   // the mask gather runs on host positions
        pos.resize((size_t)site_off[n_chr]);
        if (colate_ingest_fetch(h, 0, site_off[n_chr], pos.data(), nullptr, nullptr, nullptr)) return die("colate_ingest_fetch");
      }
    } else {
      std::vector<float> ab, ae;
      std::vector<uint32_t> meta;
      for (int c = 0; c < n_chr; c++) {
        std::cerr << "parsing CHR: " << c + 1 << " / " << n_chr << std::endl;
        int64_t n = colate_read_mut(f_mut[c].c_str(), 0, nullptr, nullptr, nullptr, nullptr);
        if (n < 0) { std::cerr << colate_last_error() << std::endl; exit(1); }
        size_t o = pos.size();
        pos.resize(o + n); ab.resize(o + n); ae.resize(o + n); meta.resize(o + n);
        if (colate_read_mut(f_mut[c].c_str(), n, pos.data() + o, ab.data() + o, ae.data() + o, meta.data() + o) < 0) {
          std::cerr << colate_last_error() << std::endl;
          exit(1);
        }
        site_off[c + 1] = (int64_t)pos.size();
      }
      if (colate_set_sites(h, n_chr, site_off.data(), pos.data(), ab.data(), ae.data(), meta.data(), 0)) return die("colate_set_sites");
    }
    ph.tick("parse .mut -> device sites");
    for (int g = 0; g < 2; g++) {
      std::vector<int64_t> first(n_chr), end(n_chr);
      colate_chr_ranges(n_chr, gh[g].n, gh[g].rc.data(), first.data(), end.data());
      if (colate_set_genome(h, g, gh[g].n, first.data(), end.data(), gh[g].bp.data(), gh[g].aaf.data(), gh[g].daf.data(), gh[g].al.data(), 0))
        return die("colate_set_genome");
      gh[g] = GenomeHost();
    }
    ph.tick("upload .colate.in x2");
    const std::vector<std::string>* masks[2] = {&f_tmask, &f_rmask};
    for (int g = 0; g < 2; g++) {
      if (masks[g]->empty()) continue;
      std::vector<uint32_t> bits(((size_t)site_off[n_chr] + 31) / 32 + 1, 0);
      for (int c = 0; c < n_chr; c++) {
        if (colate_mask_bits_from_fasta((*masks[g])[c].c_str(), site_off[c + 1] - site_off[c], pos.data() + site_off[c], site_off[c], bits.data())) {
          std::cerr << colate_last_error() << std::endl;
          exit(1);
        }
      }
      if (colate_set_mask(h, g, bits.data(), 0)) return die("colate_set_mask");
    }
    ph.tick("masks");
    // stage i
    int num_blocks = 0;
    int64_t n_used = 0;
    std::vector<double> block_stats((size_t)COLATE_MAX_BLOCKS * 4 * COLATE_NUM_AGE_BINS);
    if (colate_stage1(h, 0, 1, mt, &num_blocks, block_stats.data(), nullptr, &n_used, mt)) return die("colate_stage1");
    std::cerr << "Number of blocks: " << num_blocks << std::endl;
    ph.tick("stage i");
    // stage ii
    if (R > 0) {
      std::vector<int32_t> w((size_t)R * num_blocks);
      colate_draw_block_weights(mt, R, num_blocks, w.data());
      if (colate_stage2_bootstrap(h, R, num_blocks, w.data(), block_stats.data(), age, counts.data())) return die("colate_stage2_bootstrap");
    }
  }

  // epochs, coal.cpp:3503-3646
  std::vector<double> epochs(4096), rates_init(4096, 1.0 / 20000.0);
  int ep_null = 0, E;
  if (options.count("coal")) {
    E = colate_epochs_from_coal_file(options.get("coal").c_str(), age, epochs.data(), rates_init.data(), 4096);
    if (E < 0) return die("--coal");
    for (int e = 0; e < E; e++) std::cerr << rates_init[e] << " ";
    std::cerr << std::endl;
  } else {
    if (!options.count("bins")) { std::cerr << "Error: --bins or --coal is required." << std::endl; return 1; }
    E = colate_epochs_from_bins(options.get("bins").c_str(), age, ypg, epochs.data(), 4096, &ep_null);
    if (E < 0) { std::cerr << colate_last_error() << std::endl; exit(1); }
  }
  if (!wait_for_device()) return 1;   // (the .colate_mat cache path gets here without having touched the device)
  std::cerr << "Maximising likelihood using EM.. " << std::endl;
  std::vector<double> rates((size_t)std::max(R, 1) * E, 0.0), ll(std::max(R, 1));
  std::vector<int32_t> iters(std::max(R, 1), 0);
  if (R > 0) {
    if (colate_stage3_em(h, R, E, epochs.data(), rates_init.data(), counts.data(), 100000, rates.data(), iters.data(), ll.data()))
      return die("colate_stage3_em");
    for (int i = 0; i < R; i++) std::cerr << "Bootstrap " << i + 1 << ": Total iterations " << iters[i] << std::endl;
  }
  // .coal first: for ancient samples it zeroes rates[0..ep_null] (coal.cpp:3832-3834); the fp64 side
  // output then holds exactly the values the text was printed from
  ph.tick("stage ii + iii");
  if (colate_write_coal((out + ".coal").c_str(), R, E, epochs.data(), rates.data(), is_ancient, ep_null)) return die("write .coal");
  if (colate_write_bin((out + ".bin").c_str(), R, E, epochs.data(), rates.data(), iters.data())) return die("write .bin");
  colate_destroy(h);

  rusage usage;
  getrusage(RUSAGE_SELF, &usage);
  std::cerr << "CPU Time spent: " << usage.ru_utime.tv_sec << "." << std::setfill('0') << std::setw(6) << usage.ru_utime.tv_usec
            << "s; Max Memory usage: " << usage.ru_maxrss / 1000.0 << "Mb." << std::endl;
  std::cerr << "---------------------------------------------------------" << std::endl << std::endl;
  return 0;
}

}  // namespace

int main(int argc, char* argv[])
{
  Options options;
  if (!parse(argc, argv, options)) return 1;
  if (!options.count("mode")) {
    std::cout << "Not enough arguments supplied." << std::endl;
    help();
    return 0;
  }
  const std::string mode = options.get("mode");
  int rc = 0;
  if (mode == "mut") rc = run_mut(options);
  else {
    std::cout << "####### error #######" << std::endl;
    std::cout << "Invalid or missing mode." << std::endl;
    std::cout << "This build implements --mode mut (tmp/tmp inputs). The reference's other modes "
                 "(preprocess_mut, make_tmp, calc_depth, print_tmp, CondCoalRates) are out of scope." << std::endl;
  }
  if (options.count("help")) help();
  return rc;
}
