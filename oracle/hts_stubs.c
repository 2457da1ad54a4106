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
 *   - the bcf entry points stay aborting stubs (bcf inputs are out of scope).
 * Only the reference's struct layouts (htslib/sam.h, read where it lies) are used.
 *
 * fake BAM:  "FBAM" | int32 n_targets | n_targets x { int32 len; char name[len] } |
 *            records { int32 tid; int32 pos (0-based); uint8 mapq; uint8 reverse; int32 l_qseq;
 *                      char seq[l_qseq] (ACGTN); uint8 qual[l_qseq] } until end of file */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "htslib/sam.h"

#define STUB(name) void name(void) { fprintf(stderr, "oracle/_ref: htslib symbol " #name " is stubbed (bcf inputs are out of scope)\n"); abort(); }
STUB(bcf_destroy) STUB(bcf_get_format_values)
STUB(bcf_hdr_destroy) STUB(bcf_hdr_id2int) STUB(bcf_hdr_read) STUB(bcf_init)
STUB(bcf_is_snp) STUB(bcf_read) STUB(bcf_unpack)
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

typedef struct { FILE* f; int n_targets; char** names; } fake_bam;

static int rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

htsFile* hts_open(const char* fn, const char* mode)
{
  (void)mode;
  FILE* f = fopen(fn, "rb");
  if (!f) return NULL;
  char magic[4];
  fake_bam* b = (fake_bam*)calloc(1, sizeof *b);
  if (!rd(f, magic, 4) || memcmp(magic, "FBAM", 4) || !rd(f, &b->n_targets, 4)) { fclose(f); free(b); return NULL; }
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
