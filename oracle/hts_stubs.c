/* TEST INFRASTRUCTURE ONLY.
 * The reference's coal.cpp references htslib entry points for its bcf/bam front-ends
 * (coal.cpp:595-2069).  The vendored htslib cannot be built in this image (no lzma/bz2/curl
 * headers), so the stage-level probe library and the reference CLI built by oracle/Makefile
 * resolve those symbols here:
 *   - the SAM/BAM reader entry points bam_parser uses (include/vcf/htslib.cpp:171-575:
 *     hts_open, sam_hdr_read, bam_init1, sam_read1, hts_close, bam_destroy1) are served from
 *     a "fake BAM": a flat binary file of already aligned reads (format below), so that the
 *     UNMODIFIED parse_onebambam (coal.cpp:1799-2069) and bam_parser run on synthetic reads and
 *     pin the N3 weighting variant;
 *   - likewise the BCF reader entry points vcf_parser uses (htslib.cpp:3-56: bcf_hdr_read, bcf_init, bcf_read, bcf_unpack,
 *     bcf_get_format_values for "GT", bcf_hdr_destroy, bcf_destroy) are served from a "fake BCF" (format below), so that the
 *     unmodified parse_vcfvcf (coal.cpp:907-1228) runs on synthetic genotype records;
 *   - the remaining bcf entry points stay aborting stubs.
 * Only the reference's struct layouts (htslib/sam.h, htslib/vcf.h, read where they lie) are used.
 *
 * fake BAM:  "FBAM" | int32 n_targets | n_targets x { int32 len; char name[len] } |
 *            records { int32 tid; int32 pos (0-based); uint8 mapq; uint8 reverse; int32 l_qseq;
 *                      char seq[l_qseq] (ACGTN); uint8 qual[l_qseq] } until end of file
 *
 * fake BCF:  "FBCF" | int32 n_samples | int32 ploidy |
 *            records { int32 pos (0-based); int32 n_allele; n_allele x { int32 len; char allele[len] };
 *                      int32 gt[n_samples * ploidy] (allele index per haplotype) } until end of file */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "htslib/sam.h"
#include "htslib/vcf.h"

static void stubbed(const char* name) { fprintf(stderr, "oracle/_ref: htslib symbol %s is stubbed\n", name); abort(); }
int bcf_hdr_id2int(const bcf_hdr_t* hdr, int type, const char* id) { (void)hdr; (void)type; (void)id; stubbed("bcf_hdr_id2int"); return -1; }
int bcf_is_snp(bcf1_t* v) { (void)v; stubbed("bcf_is_snp"); return 0; }
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

/* one struct behind htsFile* for both fakes: is_bcf tells which */
typedef struct { FILE* f; int n_targets; char** names; int is_bcf, n_samples, ploidy; int32_t* gt; } fake_bam;

static int rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

htsFile* hts_open(const char* fn, const char* mode)
{
  (void)mode;
  FILE* f = fopen(fn, "rb");
  if (!f) return NULL;
  char magic[4];
  fake_bam* b = (fake_bam*)calloc(1, sizeof *b);
  if (!rd(f, magic, 4)) { fclose(f); free(b); return NULL; }
  if (!memcmp(magic, "FBCF", 4)) {
    if (!rd(f, &b->n_samples, 4) || !rd(f, &b->ploidy, 4)) { fclose(f); free(b); return NULL; }
    b->f = f;
    b->is_bcf = 1;
    b->gt = (int32_t*)calloc((size_t)b->n_samples * b->ploidy + 1, 4);
    return (htsFile*)b;
  }
  if (memcmp(magic, "FBAM", 4) || !rd(f, &b->n_targets, 4)) { fclose(f); free(b); return NULL; }
  b->f = f;
  b->names = (char**)calloc((size_t)b->n_targets, sizeof(char*));
  for (int i = 0; i < b->n_targets; i++) {
    int len = 0;
    rd(f, &len, 4);
    b->names[i] = (char*)calloc((size_t)len + 1, 1);
    rd(f, b->names[i], (size_t)len);
  }
  return (htsFile*)b;    /* opaque to bam_parser: only handed back to the functions below */
}

int hts_close(htsFile* fp)
{
  fake_bam* b = (fake_bam*)fp;
  if (!b) return 0;
  fclose(b->f);
  free(b);               /* (names stay alive: the header shares them) */
  return 0;
}

sam_hdr_t* sam_hdr_read(samFile* fp)
{
  fake_bam* b = (fake_bam*)fp;
  sam_hdr_t* h = (sam_hdr_t*)calloc(1, sizeof *h);
  h->n_targets = b->n_targets;
  h->target_name = b->names;
  return h;
}

bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) { if (b) { free(b->data); free(b); } }

static int code_of(char c)
{
  switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; }
}

int sam_read1(samFile* fp, sam_hdr_t* h, bam1_t* b)
{
  (void)h;
  fake_bam* fb = (fake_bam*)fp;
  int32_t tid, pos, l;
  uint8_t mapq, rev;
  if (!rd(fb->f, &tid, 4)) return -1;                     /* end of file */
  if (!rd(fb->f, &pos, 4) || !rd(fb->f, &mapq, 1) || !rd(fb->f, &rev, 1) || !rd(fb->f, &l, 4) || l < 0) return -2;
  const size_t need = 1 + (size_t)(l + 1) / 2 + (size_t)l;
  if (b->m_data < need) { b->data = (uint8_t*)realloc(b->data, need); b->m_data = (uint32_t)need; }
  memset(&b->core, 0, sizeof b->core);
  b->core.tid = tid;
  b->core.pos = pos;
  b->core.qual = mapq;
  b->core.flag = rev ? BAM_FREVERSE : 0;
  b->core.l_qname = 1;                                    /* "" + NUL */
  b->core.n_cigar = 0;
  b->core.l_qseq = l;
  b->data[0] = 0;
  uint8_t* s = b->data + 1;
  memset(s, 0, (size_t)(l + 1) / 2);
  char* tmp = (char*)malloc((size_t)l + 1);
  if (!rd(fb->f, tmp, (size_t)l)) { free(tmp); return -2; }
  for (int i = 0; i < l; i++) s[i >> 1] |= (uint8_t)(code_of(tmp[i]) << ((~i & 1) << 2));
  free(tmp);
  if (!rd(fb->f, s + (l + 1) / 2, (size_t)l)) return -2;
  b->l_data = (int)need;
  return (int)need + 36;                                   /* > 0, as the BAM reader's byte count (htslib.cpp:382 tests ret > 0) */
}

/* ---- fake BCF ------------------------------------------------------------------------------------------------------ */
void hts_stubs_bind(const bcf_hdr_t* h, fake_bam* f);
bcf_hdr_t* bcf_hdr_read(htsFile* fp)
{
  fake_bam* b = (fake_bam*)fp;
  bcf_hdr_t* h = (bcf_hdr_t*)calloc(1, sizeof *h);
  h->n[BCF_DT_SAMPLE] = b->n_samples;           /* bcf_hdr_nsamples(hdr), htslib.cpp:11 */
  hts_stubs_bind(h, b);
  return h;
}
void bcf_hdr_destroy(bcf_hdr_t* h) { free(h); }
bcf1_t* bcf_init(void) { return (bcf1_t*)calloc(1, sizeof(bcf1_t)); }
void bcf_destroy(bcf1_t* v)
{
  if (!v) return;
  if (v->d.allele) { for (int i = 0; v->d.allele[i]; i++) free(v->d.allele[i]); free(v->d.allele); }
  free(v);
}
int bcf_unpack(bcf1_t* b, int which) { (void)b; (void)which; return 0; }

int bcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v)
{
  (void)h;
  fake_bam* b = (fake_bam*)fp;
  int32_t pos, na;
  if (!rd(b->f, &pos, 4)) return -1;                     /* end of file: the record keeps its last content (coal.cpp:1002-1006) */
  if (!rd(b->f, &na, 4) || na < 0 || na > 64) return -2;
  if (v->d.allele) { for (int i = 0; v->d.allele[i]; i++) free(v->d.allele[i]); free(v->d.allele); }
  v->d.allele = (char**)calloc((size_t)na + 1, sizeof(char*));      /* NULL-terminated: coal.cpp:1074 scans for NULL */
  for (int i = 0; i < na; i++) {
    int32_t len = 0;
    if (!rd(b->f, &len, 4) || len < 0 || len > 4096) return -2;
    v->d.allele[i] = (char*)calloc((size_t)len + 1, 1);
    if (!rd(b->f, v->d.allele[i], (size_t)len)) return -2;
  }
  v->pos = pos;
  v->n_allele = (uint32_t)na;
  if (!rd(b->f, b->gt, (size_t)b->n_samples * b->ploidy * 4)) return -2;
  return 0;
}

/* bcf_get_format_int32(hdr, rec, "GT", &gt, &ngt_arr) (htslib.cpp:52): genotypes of the record LAST READ from this header's file.
 * The fake keeps them per file; the header does not know its file, so the genotype array travels in a side table keyed by header. */
static struct { const bcf_hdr_t* h; fake_bam* f; } g_map[64];
static int g_nmap = 0;
void hts_stubs_bind(const bcf_hdr_t* h, fake_bam* f)
{
  if (g_nmap == 64) { memmove(g_map, g_map + 32, 32 * sizeof g_map[0]); g_nmap = 32; }   /* old headers are long gone */
  g_map[g_nmap].h = h; g_map[g_nmap].f = f; g_nmap++;
}

int bcf_get_format_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type)
{
  (void)line; (void)type;
  if (strcmp(tag, "GT")) { fprintf(stderr, "oracle/_ref: fake BCF serves GT only\n"); abort(); }
  fake_bam* b = NULL;
  for (int i = g_nmap - 1; i >= 0; i--) if (g_map[i].h == hdr) { b = g_map[i].f; break; }
  if (!b) { fprintf(stderr, "oracle/_ref: fake BCF: unknown header\n"); abort(); }
  const int n = b->n_samples * b->ploidy;
  if (*ndst < n) { *dst = realloc(*dst, (size_t)n * 4); *ndst = n; }
  for (int i = 0; i < n; i++) ((int32_t*)*dst)[i] = (b->gt[i] + 1) << 1;      /* bcf_gt_allele(x) == (x >> 1) - 1 */
  return n;
}
