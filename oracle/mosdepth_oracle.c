/*
 * mosdepth_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
 *
 * A plain-C, CPU-only restatement of the depth hot path of brentp/mosdepth 0.3.14, written from the
 * algorithm description in SURVEY.md and the reference sources (cited per function as
 * mosdepth.nim:<line> / depthstat.nim:<line>).  It exists so that the sm_100a CUDA path in
 * mosdepth_b200/ can be checked bit-for-bit on the same inputs.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build or run this file.  The product
 * (libmosdepth_b200.so, the mosdepth-b200 host CLI) never links, imports or executes it.
 *
 * Parity status: PINNED.  Every numeric known answer of the reference's functional-tests.sh that
 * touches the hot path is asserted against this program by
 * tests/test_oracle_golden.py, on the reference's own fixture BAMs (copied as data under
 * tests/golden/fixtures/).  The reference binary itself cannot be built here (no Nim, no htslib).
 *
 * Third-party arithmetic restated here because it is not vendored in the reference tree:
 *   - hts-nim `cumsum` (int32 in-place inclusive prefix sum), `Record.stop` = htslib bam_endpos,
 *     `consumes` = htslib bam_cigar_type table (M=3 I=1 D=2 N=2 S=1 H=0 P=0 '='=3 X=3);
 *   - htslib BGZF framing + raw DEFLATE (done with system zlib), bam_read1 record layout, and the
 *     region iterator's overlap rule (tid match, pos < end, endpos > beg).
 *
 * Output files are the reference's (mosdepth.nim:641-682) with the *.bed.gz files written as plain
 * text `*.bed` -- parity is judged on decompressed content.
 *
 * Usage: mosdepth_oracle [options] <prefix> <bam>      (same option letters as mosdepth.nim:886-916)
 */
#define _GNU_SOURCE
#include <errno.h>
#include <inttypes.h>
#include <limits.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

static void die(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3), noreturn));
#include <stdarg.h>
static void die(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); exit(code);
}
static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) die(1, "oom"); return p; }
static void *xcalloc(size_t n, size_t m) { void *p = calloc(n ? n : 1, m ? m : 1); if (!p) die(1, "oom"); return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) die(1, "oom"); return p; }

/* ------------------------------------------------------------------------------------------------
 * BGZF reader (htslib bgzf.c behaviour restated).  The whole file is mmapped;
 * the block chain is chased once; blocks are inflated in order, optionally by `threads` workers
 * (mirrors `open(bam, threads=...)`, mosdepth.nim:888,968: only decompression is threaded).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t coff; uint32_t clen; uint32_t dlen_off; uint32_t isize; } bgzf_blk;

typedef struct {
    const uint8_t *file; uint64_t fsize;
    bgzf_blk *blk; int64_t nblk;
    /* ring of inflated slots */
    int nslot; uint8_t **slot_buf; volatile int64_t *slot_blk; /* which block is ready in the slot, -1 none */
    volatile int64_t next_claim;   /* next block index a worker will claim */
    volatile int64_t consumed;     /* main has finished blocks < consumed */
    int threads; pthread_t *th; volatile int stop;
    /* workers sleep while the ring is full, the consumer sleeps while its block is not ready (no spinning: r1 measured the
     * sched_yield() version SLOWER with 31 workers than with 15, VERDICT r1 "weak" #10) */
    pthread_mutex_t mu; pthread_cond_t cv_space, cv_ready; int space_waiters;
    /* consumer cursor */
    int64_t cur_blk; uint32_t cur_off; const uint8_t *cur_buf; uint32_t cur_len; uint32_t pending_uoff;
} bgzf_t;

static int inflate_block(const bgzf_t *bz, int64_t i, uint8_t *out) {
    const bgzf_blk *b = &bz->blk[i];
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    zs.next_in = (Bytef *)(bz->file + b->coff + b->dlen_off);
    zs.avail_in = b->clen - b->dlen_off - 8;
    zs.next_out = out; zs.avail_out = 65536;
    int r = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (r != Z_STREAM_END || zs.total_out != b->isize) return -1;
    uint32_t crc; memcpy(&crc, bz->file + b->coff + b->clen - 8, 4);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), out, b->isize) != crc) return -2;
    return 0;
}

static void *bgzf_worker(void *arg) {
    bgzf_t *bz = (bgzf_t *)arg;
    for (;;) {
        int64_t i = __sync_fetch_and_add(&bz->next_claim, 1);
        if (i >= bz->nblk || bz->stop) return NULL;
        if (i >= bz->consumed + bz->nslot) {
            pthread_mutex_lock(&bz->mu);
            bz->space_waiters++;
            while (i >= bz->consumed + bz->nslot && !bz->stop) pthread_cond_wait(&bz->cv_space, &bz->mu);
            bz->space_waiters--;
            pthread_mutex_unlock(&bz->mu);
            if (bz->stop) return NULL;
        }
        int s = (int)(i % bz->nslot);
        if (inflate_block(bz, i, bz->slot_buf[s]) != 0) die(1, "[oracle] BGZF inflate/CRC error in block %" PRId64, i);
        pthread_mutex_lock(&bz->mu);
        bz->slot_blk[s] = i;
        pthread_cond_signal(&bz->cv_ready);
        pthread_mutex_unlock(&bz->mu);
    }
}

static void bgzf_open(bgzf_t *bz, const char *path, int threads) {
    memset(bz, 0, sizeof *bz);
    int fd = open(path, O_RDONLY);
    if (fd < 0) die(1, "[oracle] cannot open %s: %s", path, strerror(errno));
    struct stat st; fstat(fd, &st); bz->fsize = (uint64_t)st.st_size;
    bz->file = (const uint8_t *)mmap(NULL, bz->fsize ? bz->fsize : 1, PROT_READ, MAP_PRIVATE, fd, 0);
    if (bz->file == MAP_FAILED) die(1, "[oracle] mmap failed");
    close(fd);
    /* chase BSIZE */
    int64_t cap = 1024; bz->blk = (bgzf_blk *)xmalloc(cap * sizeof(bgzf_blk));
    uint64_t p = 0;
    while (p + 18 <= bz->fsize) {
        const uint8_t *h = bz->file + p;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) die(1, "[oracle] not BGZF at %" PRIu64, p);
        uint32_t xlen = h[10] | (h[11] << 8), q = 12, bsize = 0; int have = 0;
        while (q + 4 <= 12 + xlen) {
            uint32_t slen = h[q + 2] | (h[q + 3] << 8);
            if (h[q] == 66 && h[q + 1] == 67 && slen == 2) { bsize = h[q + 4] | (h[q + 5] << 8); have = 1; }
            q += 4 + slen;
        }
        if (!have) die(1, "[oracle] BGZF block without BC field");
        if (bz->nblk == cap) { cap *= 2; bz->blk = (bgzf_blk *)xrealloc(bz->blk, cap * sizeof(bgzf_blk)); }
        bgzf_blk *b = &bz->blk[bz->nblk++];
        b->coff = p; b->clen = bsize + 1; b->dlen_off = 12 + xlen;
        if (p + b->clen > bz->fsize) die(1, "[oracle] truncated BGZF block");
        memcpy(&b->isize, bz->file + p + b->clen - 4, 4);
        p += b->clen;
    }
    bz->threads = threads < 0 ? 0 : threads;
    bz->nslot = bz->threads ? 8 * bz->threads + 8 : 1;
    bz->slot_buf = (uint8_t **)xmalloc(bz->nslot * sizeof(uint8_t *));
    bz->slot_blk = (volatile int64_t *)xmalloc(bz->nslot * sizeof(int64_t));
    for (int i = 0; i < bz->nslot; i++) { bz->slot_buf[i] = (uint8_t *)xmalloc(65536); bz->slot_blk[i] = -1; }
    bz->cur_blk = -1;
    pthread_mutex_init(&bz->mu, NULL); pthread_cond_init(&bz->cv_space, NULL); pthread_cond_init(&bz->cv_ready, NULL);
}

static int64_t bgzf_find_block(const bgzf_t *bz, uint64_t coff) {
    int64_t lo = 0, hi = bz->nblk - 1;
    while (lo <= hi) { int64_t m = (lo + hi) / 2; if (bz->blk[m].coff == coff) return m; if (bz->blk[m].coff < coff) lo = m + 1; else hi = m - 1; }
    return -1;
}

/* (re)start sequential reading at a virtual offset */
static void bgzf_seek(bgzf_t *bz, uint64_t voff) {
    if (bz->th) {
        pthread_mutex_lock(&bz->mu); bz->stop = 1; pthread_cond_broadcast(&bz->cv_space); pthread_mutex_unlock(&bz->mu);
        for (int i = 0; i < bz->threads; i++) pthread_join(bz->th[i], NULL);
        free(bz->th); bz->th = NULL; bz->stop = 0;
    }
    int64_t b = bgzf_find_block(bz, voff >> 16);
    if (b < 0) die(1, "[oracle] bad virtual offset");
    for (int i = 0; i < bz->nslot; i++) bz->slot_blk[i] = -1;
    bz->next_claim = b; bz->consumed = b; bz->cur_blk = b - 1; bz->cur_len = 0; bz->cur_off = 0;
    bz->pending_uoff = (uint32_t)(voff & 0xffff);
    if (bz->threads) {
        bz->th = (pthread_t *)xmalloc(bz->threads * sizeof(pthread_t));
        for (int i = 0; i < bz->threads; i++) pthread_create(&bz->th[i], NULL, bgzf_worker, bz);
    }
}

static int bgzf_next_block(bgzf_t *bz) {
    int64_t i = bz->cur_blk + 1;
    if (i >= bz->nblk) return 0;
    int s = (int)(i % bz->nslot);
    if (bz->threads) {
        pthread_mutex_lock(&bz->mu);
        if (i > bz->consumed) { bz->consumed = i; if (bz->space_waiters) pthread_cond_broadcast(&bz->cv_space); }   /* blocks < i are finished: their slots may be reused */
        while (bz->slot_blk[s] != i) pthread_cond_wait(&bz->cv_ready, &bz->mu);
        pthread_mutex_unlock(&bz->mu);
    } else if (i > bz->consumed) bz->consumed = i;
    if (bz->threads) { /* inflated by a worker */ }
    else if (inflate_block(bz, i, bz->slot_buf[s]) != 0) die(1, "[oracle] BGZF inflate/CRC error in block %" PRId64, i);
    bz->cur_blk = i; bz->cur_buf = bz->slot_buf[s]; bz->cur_len = bz->blk[i].isize;
    bz->cur_off = bz->pending_uoff <= bz->cur_len ? bz->pending_uoff : bz->cur_len; bz->pending_uoff = 0;
    return 1;
}

/* read exactly n bytes; returns n, or 0 at clean EOF (nothing read), or -1 on truncated data */
static int64_t bgzf_read(bgzf_t *bz, void *dst, int64_t n) {
    int64_t got = 0; uint8_t *d = (uint8_t *)dst;
    while (got < n) {
        if (bz->cur_off >= bz->cur_len) { if (!bgzf_next_block(bz)) return got == 0 ? 0 : -1; continue; }
        int64_t k = bz->cur_len - bz->cur_off; if (k > n - got) k = n - got;
        memcpy(d + got, bz->cur_buf + bz->cur_off, (size_t)k); got += k; bz->cur_off += (uint32_t)k;
    }
    return got;
}
static uint64_t bgzf_tell(const bgzf_t *bz) {
    if (bz->cur_blk < 0) return bz->blk[0].coff << 16;
    if (bz->cur_off >= bz->cur_len && bz->cur_blk + 1 < bz->nblk) return bz->blk[bz->cur_blk + 1].coff << 16;
    return (bz->blk[bz->cur_blk].coff << 16) | bz->cur_off;
}

/* ------------------------------------------------------------------------------------------------
 * BAM header + records ([ext] SAM/BAM spec)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { char *name; uint32_t length; int tid; } target_t;   /* hts.Target */

typedef struct {
    int32_t tid, pos; uint8_t l_qname, mapq; uint32_t n_cigar; uint16_t flag; int32_t l_seq, mtid, mpos, isize;
    uint8_t *data; uint32_t l_data, m_data;   /* qname, cigar, seq, qual, aux */
} rec_t;

#define FLAG_PROPER 2
#define FLAG_UNMAP 4
#define FLAG_READ2 128
#define FLAG_SUPP 2048

static inline const char *rec_qname(const rec_t *r) { return (const char *)r->data; }
static inline const uint32_t *rec_cigar(const rec_t *r) { return (const uint32_t *)(r->data + r->l_qname); } /* may be unaligned: use memcpy */
static inline uint32_t cig_at(const rec_t *r, int i) { uint32_t v; memcpy(&v, r->data + r->l_qname + 4 * (size_t)i, 4); return v; }
/* bam_cigar_type: bit0 = consumes query, bit1 = consumes reference (htslib sam.h, table 0x3C1A7) */
static inline int cig_type(uint32_t op) { return (0x3C1A7 >> ((op & 15) << 1)) & 3; }

/* htslib bam_endpos: pos + reference length (0 if unmapped), length 0 treated as 1 -> hts-nim Record.stop */
static int64_t rec_stop(const rec_t *r) {
    int64_t rlen = 0;
    if (!(r->flag & FLAG_UNMAP)) for (uint32_t i = 0; i < r->n_cigar; i++) { uint32_t c = cig_at(r, (int)i); if (cig_type(c) & 2) rlen += c >> 4; }
    if (rlen == 0) rlen = 1;
    return (int64_t)r->pos + rlen;
}

/* [ext] htslib sam.c bam_tag2cigar, called by bam_read1 (the reader behind hts-nim's `for rec in bam.query(...)`,
 * mosdepth.nim:185): a CIGAR with more than 65,535 operations does not fit the 16-bit n_cigar_op field, so the writer
 * stores the placeholder <l_seq>S<reflen>N and keeps the real CIGAR in a CG:B,I tag (SAM spec 4.2.2); the reader swaps
 * it back in and drops the tag.  Conditions as in htslib: n_cigar > 0, tid >= 0, pos >= 0, first op == S of length
 * l_seq, first "CG" tag is of type B with subtype I (or i), and its count is >= n_cigar and < 2^29. */
static void rec_restore_long_cigar(rec_t *r) {
    if (r->n_cigar == 0 || r->tid < 0 || r->pos < 0) return;
    uint32_t c0; memcpy(&c0, r->data + r->l_qname, 4);
    if ((c0 & 15) != 4 || (c0 >> 4) != (uint32_t)r->l_seq) return;
    const size_t fake = 4 * (size_t)r->n_cigar;
    size_t p = (size_t)r->l_qname + fake + (size_t)((r->l_seq + 1) / 2) + (size_t)r->l_seq;   /* first aux byte */
    size_t tag_at = 0, tag_end = 0; int found = 0;
    while (p + 3 <= r->l_data) {
        const uint8_t *a = r->data + p; const uint8_t t = a[2]; const size_t at = p; p += 3;
        const int match = a[0] == 'C' && a[1] == 'G';
        switch (t) {
        case 'A': case 'c': case 'C': p += 1; break;
        case 's': case 'S': p += 2; break;
        case 'i': case 'I': case 'f': p += 4; break;
        case 'd': p += 8; break;
        case 'Z': case 'H': p += strnlen((const char *)r->data + p, r->l_data - p) + 1; break;
        case 'B': { if (p + 5 > r->l_data) return; uint8_t st = r->data[p]; uint32_t cnt; memcpy(&cnt, r->data + p + 1, 4); size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; p += 5 + es * (size_t)cnt; break; }
        default: return;                                                   /* bad aux data */
        }
        if (match) { if (p > r->l_data) return; tag_at = at; tag_end = p; found = 1; break; }   /* bam_aux_get: the first CG tag */
    }
    if (!found) return;
    const uint8_t *cg = r->data + tag_at + 2;                              /* type byte */
    if (cg[0] != 'B' || !(cg[1] == 'I' || cg[1] == 'i')) return;
    uint32_t cg_len; memcpy(&cg_len, cg + 2, 4);
    if (cg_len < r->n_cigar || cg_len >= (1u << 29)) return;
    /* new layout: qname | real CIGAR | seq, qual, aux before CG | aux after CG */
    const size_t real = 4 * (size_t)cg_len, mid_beg = (size_t)r->l_qname + fake;
    const size_t new_len = (size_t)r->l_qname + real + (tag_at - mid_beg) + ((size_t)r->l_data - tag_end);
    uint8_t *nd = (uint8_t *)xmalloc(new_len + 256), *w = nd;
    memcpy(w, r->data, r->l_qname); w += r->l_qname;
    memcpy(w, cg + 6, real); w += real;
    memcpy(w, r->data + mid_beg, tag_at - mid_beg); w += tag_at - mid_beg;
    memcpy(w, r->data + tag_end, (size_t)r->l_data - tag_end);
    free(r->data); r->data = nd; r->l_data = (uint32_t)new_len; r->m_data = (uint32_t)new_len + 256; r->n_cigar = cg_len;
}

static int bam_read_record(bgzf_t *bz, rec_t *r) {
    int32_t bs; int64_t k = bgzf_read(bz, &bs, 4);
    if (k == 0) return 0;
    if (k < 0 || bs < 32) die(1, "[oracle] truncated/corrupt BAM record");
    uint8_t core[32]; if (bgzf_read(bz, core, 32) != 32) die(1, "[oracle] truncated BAM record");
    memcpy(&r->tid, core, 4); memcpy(&r->pos, core + 4, 4); r->l_qname = core[8]; r->mapq = core[9];
    { uint16_t nc; memcpy(&nc, core + 12, 2); r->n_cigar = nc; } memcpy(&r->flag, core + 14, 2); memcpy(&r->l_seq, core + 16, 4);
    memcpy(&r->mtid, core + 20, 4); memcpy(&r->mpos, core + 24, 4); memcpy(&r->isize, core + 28, 4);
    r->l_data = (uint32_t)bs - 32;
    if (r->l_data > r->m_data) { r->m_data = r->l_data + 256; r->data = (uint8_t *)xrealloc(r->data, r->m_data); }
    if (r->l_data && bgzf_read(bz, r->data, r->l_data) != (int64_t)r->l_data) die(1, "[oracle] truncated BAM record");
    rec_restore_long_cigar(r);
    return 1;
}

/* aux tag lookup: tag[string](rec, "RG") (mosdepth.nim:283). returns pointer to Z string or NULL */
static const char *rec_aux_z(const rec_t *r, const char tag[2]) {
    size_t p = (size_t)r->l_qname + 4 * (size_t)r->n_cigar + (size_t)((r->l_seq + 1) / 2) + (size_t)r->l_seq;
    while (p + 3 <= r->l_data) {
        const uint8_t *a = r->data + p; uint8_t t = a[2]; p += 3;
        int match = a[0] == (uint8_t)tag[0] && a[1] == (uint8_t)tag[1];
        switch (t) {
        case 'A': case 'c': case 'C': p += 1; break;
        case 's': case 'S': p += 2; break;
        case 'i': case 'I': case 'f': p += 4; break;
        case 'd': p += 8; break;
        case 'Z': case 'H': { const char *s = (const char *)r->data + p; size_t n = strnlen(s, r->l_data - p); if (match && t == 'Z') return s; p += n + 1; break; }
        case 'B': { uint8_t st = r->data[p]; int32_t cnt; memcpy(&cnt, r->data + p + 1, 4); size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4; p += 5 + es * (size_t)cnt; break; }
        default: return NULL;
        }
        if (match) return NULL; /* tag present but not a string */
    }
    return NULL;
}

typedef struct { bgzf_t bz; target_t *targets; int n_targets; rec_t look; int have_look; int eof; uint64_t first_rec_voff; const char *path; } bam_t;

static void bam_open(bam_t *b, const char *path, int threads) {
    memset(b, 0, sizeof *b); b->path = path;
    bgzf_open(&b->bz, path, threads);
    if (b->bz.nblk == 0) die(1, "[oracle] empty file");
    bgzf_seek(&b->bz, b->bz.blk[0].coff << 16);
    char magic[4]; int32_t l_text, n_ref;
    if (bgzf_read(&b->bz, magic, 4) != 4 || memcmp(magic, "BAM\1", 4)) die(1, "[oracle] not a BAM file");
    bgzf_read(&b->bz, &l_text, 4);
    char *text = (char *)xmalloc((size_t)l_text + 1); if (l_text) bgzf_read(&b->bz, text, l_text); free(text);
    bgzf_read(&b->bz, &n_ref, 4);
    b->n_targets = n_ref; b->targets = (target_t *)xcalloc((size_t)n_ref, sizeof(target_t));
    for (int i = 0; i < n_ref; i++) {
        int32_t l_name; bgzf_read(&b->bz, &l_name, 4);
        b->targets[i].name = (char *)xmalloc((size_t)l_name + 1); bgzf_read(&b->bz, b->targets[i].name, l_name); b->targets[i].name[l_name] = 0;
        int32_t l_ref; bgzf_read(&b->bz, &l_ref, 4); b->targets[i].length = (uint32_t)l_ref; b->targets[i].tid = i;
    }
    b->first_rec_voff = bgzf_tell(&b->bz);
}

/* ------------------------------------------------------------------------------------------------
 * Index: only what is needed to emulate `bam.query(tid, 0, len)` cheaply for `-c`: the smallest chunk
 * start of the contig (BAI / CSI pseudo-bin or min over bins).  Without an index file the oracle
 * refuses to run, like mosdepth.nim:969-971.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int n_ref; uint64_t *beg; uint8_t *has; } index_t;

static uint8_t *slurp_maybe_bgzf(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb"); if (!f) return NULL;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *raw = (uint8_t *)xmalloc((size_t)sz + 1); if (fread(raw, 1, (size_t)sz, f) != (size_t)sz) die(1, "read error"); fclose(f);
    if (sz >= 18 && raw[0] == 0x1f && raw[1] == 0x8b) { /* gzip/BGZF members, concatenate */
        size_t cap = (size_t)sz * 8 + 65536, len = 0; uint8_t *out = (uint8_t *)xmalloc(cap); size_t p = 0;
        while (p < (size_t)sz) {
            z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, 15 + 16);
            zs.next_in = raw + p; zs.avail_in = (uInt)((size_t)sz - p);
            int r;
            do { if (cap - len < 65536) { cap *= 2; out = (uint8_t *)xrealloc(out, cap); }
                 zs.next_out = out + len; zs.avail_out = (uInt)(cap - len); r = inflate(&zs, Z_NO_FLUSH); len = cap - zs.avail_out - 0; len = (size_t)(zs.next_out - out);
            } while (r == Z_OK);
            if (r != Z_STREAM_END) die(1, "[oracle] bad gzip index");
            p = (size_t)sz - zs.avail_in; inflateEnd(&zs);
        }
        free(raw); *n = len; return out;
    }
    *n = (size_t)sz; return raw;
}

static int index_load(index_t *ix, const char *bam_path) {
    memset(ix, 0, sizeof *ix);
    char p[4096]; size_t n = 0; uint8_t *d = NULL; int is_csi = 0;
    snprintf(p, sizeof p, "%s.csi", bam_path); d = slurp_maybe_bgzf(p, &n); if (d) is_csi = 1;
    if (!d) { snprintf(p, sizeof p, "%s.bai", bam_path); d = slurp_maybe_bgzf(p, &n); }
    if (!d) { size_t L = strlen(bam_path); if (L > 4) { snprintf(p, sizeof p, "%.*s.bai", (int)(L - 4), bam_path); d = slurp_maybe_bgzf(p, &n); } }
    if (!d) return -1;
    size_t q = 0;
    uint32_t pseudo_bin = 37450;
    if (is_csi) { if (n < 16 || memcmp(d, "CSI\1", 4)) die(1, "[oracle] bad CSI"); int32_t depth, l_aux; memcpy(&depth, d + 8, 4); memcpy(&l_aux, d + 12, 4); q = 16 + (size_t)l_aux;
        pseudo_bin = (uint32_t)(((1u << ((depth + 1) * 3)) - 1) / 7 + 1); }
    else { if (n < 8 || memcmp(d, "BAI\1", 4)) die(1, "[oracle] bad BAI"); q = 4; }
    int32_t n_ref; memcpy(&n_ref, d + q, 4); q += 4;
    ix->n_ref = n_ref; ix->beg = (uint64_t *)xcalloc((size_t)n_ref, 8); ix->has = (uint8_t *)xcalloc((size_t)n_ref, 1);
    for (int r = 0; r < n_ref; r++) {
        int32_t n_bin; memcpy(&n_bin, d + q, 4); q += 4; uint64_t mn = UINT64_MAX;
        for (int b = 0; b < n_bin; b++) {
            uint32_t bin; memcpy(&bin, d + q, 4); q += 4; if (is_csi) q += 8;
            int32_t n_chunk; memcpy(&n_chunk, d + q, 4); q += 4;
            int pseudo = (bin == pseudo_bin);
            for (int c = 0; c < n_chunk; c++) { uint64_t cb; memcpy(&cb, d + q, 8); q += 16; if (!pseudo && cb < mn) mn = cb; }
        }
        if (!is_csi) { int32_t n_intv; memcpy(&n_intv, d + q, 4); q += 4 + 8 * (size_t)n_intv; }
        if (n_bin > 0 && mn != UINT64_MAX) { ix->has[r] = 1; ix->beg[r] = mn; }
    }
    free(d); return 0;
}

/* ------------------------------------------------------------------------------------------------
 * depthstat.nim
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t cum_length; uint64_t cum_depth; uint32_t min_depth, max_depth; } depth_stat; /* depthstat.nim:2-6 */
static void ds_clear(depth_stat *d) { d->cum_length = 0; d->cum_depth = 0; d->min_depth = UINT32_MAX; d->max_depth = 0; } /* :47-51 */
static depth_stat ds_new(const int32_t *v, int64_t n) { /* newDepthStat depthstat.nim:39-45 */
    depth_stat r; r.cum_length = n; r.cum_depth = 0; r.min_depth = UINT32_MAX; r.max_depth = 0;
    for (int64_t i = 0; i < n; i++) { r.cum_depth += (uint64_t)(int64_t)v[i]; uint32_t u = (uint32_t)v[i]; if (u < r.min_depth) r.min_depth = u; if (u > r.max_depth) r.max_depth = u; }
    return r;
}
static depth_stat ds_add(depth_stat a, depth_stat b) { /* `+` depthstat.nim:53-57 */
    depth_stat r; r.cum_length = a.cum_length + b.cum_length; r.cum_depth = a.cum_depth + b.cum_depth;
    r.min_depth = a.min_depth < b.min_depth ? a.min_depth : b.min_depth; r.max_depth = a.max_depth > b.max_depth ? a.max_depth : b.max_depth; return r;
}
typedef struct { uint32_t *counts; int64_t len; int64_t n; } count_stat;  /* CountStat depthstat.nim:9-37 */
static void cs_clear(count_stat *c) { if (c->n == 0) return; c->n = 0; memset(c->counts, 0, (size_t)c->len * 4); }
static void cs_add(count_stat *c, int32_t value) {
    c->n++;
    if (value < 0) die(1, "error setting negative depth value:%d", value);   /* IndexDefect depthstat.nim:18-19 */
    int64_t hi = c->len - 1; c->counts[(int64_t)value < hi ? (int64_t)value : hi]++;
}
static int64_t cs_median(const count_stat *c) {   /* depthstat.nim:23-30 */
    int64_t stop_n = (int64_t)(0.5 + (double)c->n * 0.5), cum = 0;
    for (int64_t i = 0; i < c->len; i++) { cum += c->counts[i]; if (cum >= stop_n) return i; }
    return -1;
}

/* ------------------------------------------------------------------------------------------------
 * CIGAR -> start/end events: gen_start_ends (mosdepth.nim:146-168), inc_coverage (:170-172)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t pos; int32_t value; } pair_t;
typedef struct { pair_t *a; int n, m; } pairvec;
static void pv_push(pairvec *v, int64_t pos, int32_t val) { if (v->n == v->m) { v->m = v->m ? 2 * v->m : 64; v->a = (pair_t *)xrealloc(v->a, (size_t)v->m * sizeof(pair_t)); } v->a[v->n].pos = pos; v->a[v->n].value = val; v->n++; }

static void gen_start_ends(const rec_t *r, int64_t ipos, pairvec *out) {
    if (r->n_cigar == 1 && (cig_at(r, 0) & 15) == 0) {            /* :148-150 single M */
        pv_push(out, ipos, 1); pv_push(out, ipos + (cig_at(r, 0) >> 4), -1); return;
    }
    int64_t pos = ipos, last_stop = -1;
    for (uint32_t i = 0; i < r->n_cigar; i++) {                   /* :155-166 */
        uint32_t c = cig_at(r, i); int con = cig_type(c);
        if (!(con & 2)) continue;
        int64_t olen = c >> 4;
        if (con & 1) {
            if (pos != last_stop) { pv_push(out, pos, 1); if (last_stop != -1) pv_push(out, last_stop, -1); }
            last_stop = pos + olen;
        }
        pos += olen;
    }
    if (last_stop != -1) pv_push(out, last_stop, -1);             /* :167-168 */
}

static int32_t *g_arr; static int64_t g_arr_len, g_arr_cap;   /* coverage_t (mosdepth.nim:44) */
static inline void arr_add(int64_t pos, int32_t v) {
    /* the reference indexes unchecked (release build); out-of-range events only arise from invalid BAMs
     * (read past contig end).  They are dropped here and in the CUDA path alike (DESIGN.md "Divergences"). */
    if (pos < 0 || pos >= g_arr_len) return;
    g_arr[pos] = (int32_t)((uint32_t)g_arr[pos] + (uint32_t)v);
}
static void arr_init(int64_t tlen) {   /* init mosdepth.nim:238-248 */
    if (tlen > g_arr_cap) { g_arr_cap = tlen; free(g_arr); g_arr = (int32_t *)xmalloc((size_t)tlen * 4); }
    g_arr_len = tlen; memset(g_arr, 0, (size_t)tlen * 4);
}

/* seen: Table[string, Record] (mosdepth.nim:262) -- chained hash keyed by qname */
typedef struct seen_ent { struct seen_ent *next; uint64_t h; rec_t rec; } seen_ent;
typedef struct { seen_ent **b; size_t nb, n; seen_ent *freelist; } seen_t;
static uint64_t hash_str(const char *s) { uint64_t h = 1469598103934665603ULL; while (*s) { h ^= (uint8_t)*s++; h *= 1099511628211ULL; } return h; }
static void seen_init(seen_t *t) { memset(t, 0, sizeof *t); t->nb = 1 << 16; t->b = (seen_ent **)xcalloc(t->nb, sizeof(seen_ent *)); }
static void seen_grow(seen_t *t) {
    size_t nb = t->nb * 4; seen_ent **b = (seen_ent **)xcalloc(nb, sizeof(seen_ent *));
    for (size_t i = 0; i < t->nb; i++) for (seen_ent *e = t->b[i], *nx; e; e = nx) { nx = e->next; size_t k = e->h & (nb - 1); e->next = b[k]; b[k] = e; }
    free(t->b); t->b = b; t->nb = nb;
}
static seen_ent *seen_find(seen_t *t, const char *q, uint64_t h) { for (seen_ent *e = t->b[h & (t->nb - 1)]; e; e = e->next) if (e->h == h && !strcmp(rec_qname(&e->rec), q)) return e; return NULL; }
static void rec_copy(rec_t *d, const rec_t *s) { uint8_t *data = d->data; uint32_t m = d->m_data; *d = *s; d->data = data; d->m_data = m; if (s->l_data > m) { d->m_data = s->l_data; d->data = (uint8_t *)xrealloc(d->data, d->m_data); } memcpy(d->data, s->data, s->l_data); }
static void seen_put(seen_t *t, const rec_t *r) {  /* seen[rc.qname] = rc  (overwrite) */
    uint64_t h = hash_str(rec_qname(r)); seen_ent *e = seen_find(t, rec_qname(r), h);
    if (!e) { if (t->n > t->nb) seen_grow(t); if (t->freelist) { e = t->freelist; t->freelist = e->next; } else e = (seen_ent *)xcalloc(1, sizeof(seen_ent));
              e->h = h; size_t k = h & (t->nb - 1); e->next = t->b[k]; t->b[k] = e; t->n++; }
    rec_copy(&e->rec, r);
}
static int seen_take(seen_t *t, const char *q, rec_t **mate_out, seen_ent **ent_out) { /* seen.take(qname, mate) */
    uint64_t h = hash_str(q); seen_ent **pp = &t->b[h & (t->nb - 1)];
    for (; *pp; pp = &(*pp)->next) if ((*pp)->h == h && !strcmp(rec_qname(&(*pp)->rec), q)) { seen_ent *e = *pp; *pp = e->next; t->n--; *mate_out = &e->rec; *ent_out = e; return 1; }
    return 0;
}
static void seen_release(seen_t *t, seen_ent *e) { e->next = t->freelist; t->freelist = e; }
static void seen_clear(seen_t *t) { for (size_t i = 0; i < t->nb; i++) { for (seen_ent *e = t->b[i], *nx; e; e = nx) { nx = e->next; seen_release(t, e); } t->b[i] = NULL; } t->n = 0; }

static int pair_cmp(const void *a, const void *b) { const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b; return x->pos < y->pos ? -1 : x->pos > y->pos; }
/* stable merge sort is what Nim's algorithm.sort is; ties between +1/-1 at one position keep insertion order */
static void stable_sort_pairs(pair_t *a, int n) {
    if (n < 2) return; pair_t *tmp = (pair_t *)xmalloc((size_t)n * sizeof(pair_t));
    for (int w = 1; w < n; w *= 2) { for (int lo = 0; lo < n; lo += 2 * w) { int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = pair_cmp(&a[j], &a[i]) < 0 ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++]; while (j < hi) tmp[k++] = a[j++]; }
        memcpy(a, tmp, (size_t)n * sizeof(pair_t)); }
    free(tmp);
}

/* options shared by coverage() */
typedef struct {
    int mapq; int64_t min_len, max_len; uint16_t eflag, iflag; char **read_groups; int n_rg; int fast_mode, fragment_mode;
} cov_opts;

static int64_t g_records_total = 0;  /* all records yielded by the per-contig queries (for reads/s reporting) */

/* coverage() mosdepth.nim:250-360.  Records of one contig are pulled from the sequential reader:
 * equivalent to bam.query(tid, 0, len) on a coordinate-sorted, indexed BAM (overlap rule pos < len). */
static int coverage(bam_t *bam, int tid, const cov_opts *o, seen_t *seen) {
    target_t *tgt = &bam->targets[tid];
    int found = 0; pairvec ses = {0, 0, 0};
    seen_clear(seen);
    for (;;) {
        if (!bam->have_look) { if (bam->eof || !bam_read_record(&bam->bz, &bam->look)) { bam->eof = 1; break; } bam->have_look = 1; }
        rec_t *rec = &bam->look;
        if (rec->tid < 0) { bam->eof = 1; bam->have_look = 0; break; }      /* unplaced tail: never returned by a tid query */
        if (rec->tid < tid) { bam->have_look = 0; continue; }                 /* contig skipped by the caller */
        if (rec->tid > tid) break;
        bam->have_look = 0;
        if ((uint32_t)rec->pos >= tgt->length || rec->pos < 0) continue;      /* iterator overlap rule: pos < end */
        g_records_total++;
        if (!found) { arr_init((int64_t)tgt->length + 1); found = 1; }       /* :273-275 */
        if ((int)rec->mapq < o->mapq) continue;                                 /* :276 */
        int64_t ais = rec->isize < 0 ? -(int64_t)rec->isize : rec->isize;
        if (ais < o->min_len || ais > o->max_len) continue;                     /* :277 */
        if ((rec->flag & o->eflag) != 0) continue;                              /* :278-279 */
        if (o->iflag != 0 && (rec->flag & o->iflag) == 0) continue;             /* :280-281 */
        if (o->n_rg > 0) {                                                      /* :282-285 */
            const char *t = rec_aux_z(rec, "RG"); int ok = 0;
            if (t) for (int i = 0; i < o->n_rg; i++) if (!strcmp(t, o->read_groups[i])) ok = 1;
            if (!ok) continue;
        }
        int64_t start = rec->pos;
        /* :292-340 mate overlap */
        if (!o->fast_mode && !o->fragment_mode && (rec->flag & FLAG_PROPER) && !(rec->flag & FLAG_SUPP)) {
            int64_t stop = rec_stop(rec);
            int haskey = 0;
            if (rec->tid == rec->mtid && stop > rec->mpos && start == rec->mpos) haskey = seen_find(seen, rec_qname(rec), hash_str(rec_qname(rec))) != NULL;
            if (rec->tid == rec->mtid && stop > rec->mpos && (start < rec->mpos || (start == rec->mpos && !haskey))) {
                seen_put(seen, rec);                                            /* :299-300 */
            } else {
                rec_t *mate; seen_ent *ent;
                if (seen_take(seen, rec_qname(rec), &mate, &ent)) {             /* :302 */
                    if (rec->n_cigar == 1 && mate->n_cigar == 1) {              /* :307-309 */
                        arr_add(start, -1); arr_add(rec_stop(mate), 1);
                    } else {                                                    /* :323-338 */
                        ses.n = 0; gen_start_ends(rec, start, &ses); gen_start_ends(mate, mate->pos, &ses);
                        stable_sort_pairs(ses.a, ses.n);
                        int pair_depth = 0; int64_t last_pos = 0;
                        for (int i = 0; i < ses.n; i++) {
                            if (ses.a[i].value == -1 && pair_depth == 2) { arr_add(last_pos, -1); arr_add(ses.a[i].pos, 1); }
                            pair_depth += ses.a[i].value; last_pos = ses.a[i].pos;
                        }
                    }
                    seen_release(seen, ent);
                }
            }
        }
        if (o->fast_mode) {                                                     /* :342-344 */
            arr_add(start, 1); arr_add(rec_stop(rec), -1);
        } else if (o->fragment_mode) {                                          /* :346-353 */
            if ((rec->flag & FLAG_READ2) || !(rec->flag & FLAG_PROPER) || (rec->flag & FLAG_SUPP)) continue;
            int64_t fs = start < rec->mpos ? start : rec->mpos;
            arr_add(fs, 1); arr_add(fs + ais, -1);
        } else {                                                                /* :355-356 */
            ses.n = 0; gen_start_ends(rec, start, &ses);
            for (int i = 0; i < ses.n; i++) arr_add(ses.a[i].pos, ses.a[i].value);
        }
    }
    free(ses.a);
    return found ? tid : -2;                                                    /* :358-360 */
}

static void to_coverage(void) {   /* mosdepth.nim:54-56 -> hts-nim cumsum: in-place int32 inclusive prefix sum */
    uint32_t s = 0; for (int64_t i = 0; i < g_arr_len; i++) { s += (uint32_t)g_arr[i]; g_arr[i] = (int32_t)s; }
}

/* ------------------------------------------------------------------------------------------------
 * regions: window_gen :382-388, bed_to_table :362-380, bed_line_to_region :195-209
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t start, stop; char *name; } region_t;
typedef struct { char *chrom; region_t *r; int n, m; int used; } bed_chrom;
typedef struct { bed_chrom *c; int n, m; } bed_table;

static int64_t nim_parse_int(const char *s) {   /* strutils.parseInt: whole string must be an integer */
    char *e; errno = 0; if (!*s) die(1, "invalid integer: %s", s);
    long long v = strtoll(s, &e, 10); if (*e || errno) die(1, "invalid integer: %s", s); return v;
}
static char *strip_dup(const char *s) { while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n' || *s == '\v' || *s == '\f') s++; size_t n = strlen(s); while (n && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\r' || s[n - 1] == '\n' || s[n - 1] == '\v' || s[n - 1] == '\f')) n--; char *r = (char *)xmalloc(n + 1); memcpy(r, s, n); r[n] = 0; return r; }

static bed_table *bed_to_table(const char *path) {
    FILE *f = fopen(path, "r"); if (!f) die(1, "[oracle] cannot open bed %s", path);
    bed_table *t = (bed_table *)xcalloc(1, sizeof *t); char *line = NULL; size_t cap = 0; ssize_t n;
    while ((n = getline(&line, &cap, f)) > 0) {
        while (n && (line[n - 1] == '\n')) line[--n] = 0;
        if (n && line[n - 1] == '\r') line[--n] = 0;                   /* hts_getline strips \r\n */
        if (n == 0) break;                                              /* hts_getline(...) > 0 ends the loop on an empty line (:366) */
        if (line[0] == 't' && !strncmp(line, "track ", 6)) continue;   /* :367-368 */
        if (line[0] == '#') continue;                                  /* :369-370 */
        /* split('\t', 5) */
        char *fld[6]; int nf = 0; char *p = line; fld[nf++] = p;
        while (nf < 6 && (p = strchr(p, '\t'))) { *p++ = 0; fld[nf++] = p; }
        if (nf < 3) { /* restore tabs for the message */ for (int i = 1; i < nf; i++) fld[i][-1] = '\t'; char *s = strip_dup(line); fprintf(stderr, "[mosdepth] skipping bad bed line:%s\n", s); free(s); continue; }
        int64_t s = nim_parse_int(fld[1]), e = nim_parse_int(fld[2]);
        if (s > e) die(1, "[slivar] ERROR: start > end in bed line:%s", line);
        bed_chrom *bc = NULL; for (int i = 0; i < t->n; i++) if (!strcmp(t->c[i].chrom, fld[0])) { bc = &t->c[i]; break; }
        if (!bc) { if (t->n == t->m) { t->m = t->m ? 2 * t->m : 16; t->c = (bed_chrom *)xrealloc(t->c, (size_t)t->m * sizeof(bed_chrom)); } bc = &t->c[t->n++]; memset(bc, 0, sizeof *bc); bc->chrom = strdup(fld[0]); }
        if (bc->n == bc->m) { bc->m = bc->m ? 2 * bc->m : 64; bc->r = (region_t *)xrealloc(bc->r, (size_t)bc->m * sizeof(region_t)); }
        region_t *r = &bc->r[bc->n++]; r->start = (uint32_t)s; r->stop = (uint32_t)e; r->name = nf > 3 ? strip_dup(fld[3]) : strdup("");
    }
    free(line); fclose(f);
    /* stable sort by start (:376-377) */
    for (int c = 0; c < t->n; c++) { bed_chrom *bc = &t->c[c]; int nn = bc->n; if (nn < 2) continue; region_t *tmp = (region_t *)xmalloc((size_t)nn * sizeof(region_t)); region_t *a = bc->r;
        for (int w = 1; w < nn; w *= 2) { for (int lo = 0; lo < nn; lo += 2 * w) { int mid = lo + w < nn ? lo + w : nn, hi = lo + 2 * w < nn ? lo + 2 * w : nn, i = lo, j = mid, k = lo;
                while (i < mid && j < hi) tmp[k++] = ((int64_t)a[j].start - (int64_t)a[i].start) < 0 ? a[j++] : a[i++];
                while (i < mid) tmp[k++] = a[i++]; while (j < hi) tmp[k++] = a[j++]; } memcpy(a, tmp, (size_t)nn * sizeof(region_t)); }
        free(tmp); }
    return t;
}
static bed_chrom *bed_find(bed_table *t, const char *chrom) { if (!t) return NULL; for (int i = 0; i < t->n; i++) if (!t->c[i].used && !strcmp(t->c[i].chrom, chrom)) return &t->c[i]; return NULL; }

/* imean mosdepth.nim:399-413 */
static double imean(uint32_t start, uint32_t stop, count_stat *ms) {
    uint32_t len = (uint32_t)g_arr_len;
    if (start > len) return 0;
    uint32_t e = stop < len ? stop : len;
    if (ms->len != 0) { cs_clear(ms); for (uint32_t i = start; i < e; i++) cs_add(ms, g_arr[i]); return (double)cs_median(ms); }
    double L = (double)(uint32_t)(stop - start), result = 0;
    for (uint32_t i = start; i < e; i++) result += (double)g_arr[i] / L;
    return result;
}

#define MAX_COVERAGE 400000
typedef struct { int64_t *d; int64_t len; } dist_t;
static void dist_setlen(dist_t *d, int64_t n) { if (n > d->len) { d->d = (int64_t *)xrealloc(d->d, (size_t)n * 8); memset(d->d + d->len, 0, (size_t)(n - d->len) * 8); } d->len = n; }
static void dist_new(dist_t *d, int64_t n) { d->d = (int64_t *)xcalloc((size_t)n, 8); d->len = n; }
/* inc mosdepth.nim:417-437 */
static void dist_inc(dist_t *d, uint32_t start, uint32_t stop) {
    int32_t L = (int32_t)(d->len - 1);
    if ((int64_t)start >= g_arr_len) { fprintf(stderr, "[mosdepth] warning requested interval outside of chromosome range:%u..%u\n", start, stop); return; }
    uint32_t istop = stop < (uint32_t)g_arr_len ? stop : (uint32_t)g_arr_len;
    for (uint32_t i = start; i < istop; i++) {
        int32_t v = g_arr[i];
        if (v > MAX_COVERAGE) v = MAX_COVERAGE - 10;
        if (v >= L) { dist_setlen(d, (int64_t)v + 10); L = (int32_t)(d->len - 1); }
        if (v < 0) continue;
        d->d[v] += 1;
    }
}
static void sum_into(const dist_t *from, dist_t *to) { if (from->len > to->len) dist_setlen(to, from->len); for (int64_t i = 0; i < from->len; i++) to->d[i] += from->d[i]; } /* :491-499 */

static int g_precision = 2;
/* write_distribution mosdepth.nim:439-458 */
static void write_distribution(const char *chrom, const dist_t *d, FILE *fh) {
    int64_t sum = 0; for (int64_t i = 0; i < d->len; i++) sum += d->d[i];
    if (sum < 1) return;
    double cum = 0;
    for (int64_t i = 0; i < d->len; i++) {
        int64_t irev = d->len - i - 1, v = d->d[irev];
        if (irev > 300 && v == 0) continue;
        cum += (double)v / (double)sum;
        if (cum < 8e-5) continue;
        fprintf(fh, "%s\t%" PRId64 "\t%.*f\n", chrom, irev, g_precision, cum);
    }
}
static int g_summary_header = 1;
/* write_summary mosdepth.nim:460-480 */
static void write_summary(const char *region, depth_stat st, FILE *fh) {
    double mean = st.cum_length > 0 ? (double)st.cum_depth / (double)st.cum_length : 0.0;
    uint32_t mn = st.min_depth == UINT32_MAX ? 0 : st.min_depth;
    if (g_summary_header) { fputs("chrom\tlength\tbases\tmean\tmin\tmax\n", fh); g_summary_header = 0; }
    fprintf(fh, "%s\t%" PRId64 "\t%" PRIu64 "\t%.*f\t%u\t%u\n", region, st.cum_length, st.cum_depth, g_precision, mean, mn, st.max_depth);
}

/* quantize: get_quantize_args :501-519, linear_search :98-109, make_lookup :111-122, gen_quantized :124-141 */
static int i64_cmp(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return x < y ? -1 : x > y; }
static int get_quantize_args(const char *qa, int64_t **out) {
    if (!qa) { *out = NULL; return 0; }
    size_t n = strlen(qa); char *a = (char *)xmalloc(n + 64); strcpy(a, qa);
    if (!strchr(a, ':')) { memmove(a + 1, a, n + 1); a[0] = ':'; strcat(a, ":"); }
    if (a[0] == ':') { memmove(a + 1, a, strlen(a) + 1); a[0] = '0'; }
    if (a[strlen(a) - 1] == ':') sprintf(a + strlen(a), "%lld", (long long)INT64_MAX);
    int cnt = 1; for (char *p = a; *p; p++) if (*p == ':') cnt++;
    int64_t *q = (int64_t *)xmalloc((size_t)cnt * 8); int k = 0; char *save, *tok = a;
    for (;;) { save = strchr(tok, ':'); if (save) *save = 0; char *e; errno = 0; if (!*tok) { fprintf(stderr, "[mosdepth] invalid quantize string: '%s'\n", qa); exit(2); }
        long long v = strtoll(tok, &e, 10); if (*e || errno) { fprintf(stderr, "[mosdepth] invalid quantize string: '%s'\n", qa); exit(2); } q[k++] = v; if (!save) break; tok = save + 1; }
    qsort(q, (size_t)k, 8, i64_cmp); free(a); *out = q; return k;
}
static void linear_search(int64_t q, const int64_t *vals, int n, int64_t *idx) {
    if (q < vals[0] || q > vals[n - 1]) { *idx = -1; return; }
    for (int i = 0; i < n; i++) { if (vals[i] > q) { *idx = i - 1; return; } if (vals[i] == q) { *idx = i; return; } }
    *idx = n - 1;
}
static char **make_lookup(const int64_t *quants, int nq) {
    int L = nq - 1; char **lk = (char **)xcalloc((size_t)(L > 0 ? L : 1), sizeof(char *));
    for (int i = 0; i < L; i++) { char env[64], buf[96]; snprintf(env, sizeof env, "MOSDEPTH_Q%d", i); const char *t = getenv(env);
        if (!t || !*t) { if (quants[i + 1] == INT64_MAX) snprintf(buf, sizeof buf, "%" PRId64 ":inf", quants[i]); else snprintf(buf, sizeof buf, "%" PRId64 ":%" PRId64, quants[i], quants[i + 1]); lk[i] = strdup(buf); }
        else lk[i] = strdup(t); }
    return lk;
}
static void write_quantized(FILE *fh, const char *chrom, const int64_t *quants, int nq, char **lookup) {
    if (g_arr_len <= 0) return;
    int nl = nq - 1; int64_t last_q, q; linear_search(g_arr[0], quants, nq, &last_q);
    int64_t last_pos = 0, high = g_arr_len - 1;
    for (int64_t pos = 0; pos < high - 1; pos++) {                     /* 0..<(arr.high-1)  (:132) */
        linear_search(g_arr[pos], quants, nq, &q);
        if (q == last_q) continue;
        if (last_q != -1 && last_q < nl) fprintf(fh, "%s\t%" PRId64 "\t%" PRId64 "\t%s\n", chrom, last_pos, pos, lookup[last_q]);
        last_q = q; last_pos = pos;
    }
    if (last_q != -1 && last_pos < high && last_q < nl) fprintf(fh, "%s\t%" PRId64 "\t%" PRId64 "\t%s\n", chrom, last_pos, g_arr_len - 1, lookup[last_q]);
}
/* gen_depths mosdepth.nim:58-96, consumed at :785-793 (offset 0, istop 0) */
static void write_per_base(FILE *fh, const char *chrom) {
    int64_t last_depth = -1, i = 0, last_i = 0, stop = g_arr_len - 1;
    for (int64_t k = 0; k < g_arr_len; k++) {
        int64_t depth = g_arr[k];
        if (i == stop) break;
        if (depth == last_depth) { i++; continue; }
        if (last_depth != -1) fprintf(fh, "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\n", chrom, last_i, i, last_depth);
        last_depth = depth; last_i = i;
        if (i + 1 == stop) break;
        i++;
    }
    if (last_i < stop) fprintf(fh, "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\n", chrom, last_i, g_arr_len - 1, last_depth);
    else if (last_i != i) fprintf(fh, "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\n", chrom, last_i - 1, i, last_depth);
    else fprintf(fh, "%s\t%" PRId64 "\t%" PRId64 "\t%" PRId64 "\n", chrom, last_i, i, last_depth);
}
/* write_thresholds mosdepth.nim:522-557 */
static void write_thresholds(FILE *fh, int tid, const char *chrom, const int64_t *thr, int nthr, const region_t *r) {
    if (nthr == 0) return;
    fprintf(fh, "%s\t%u\t%u\t%s", chrom, r->start, r->stop, (r->name && *r->name) ? r->name : "unknown");
    if (tid == -2) { for (int i = 0; i < nthr; i++) fputs("\t0", fh); fputc('\n', fh); return; }
    int64_t *counts = (int64_t *)xcalloc((size_t)nthr, 8);
    /* arr[start..<stop]: the reference raises on out-of-range; clamped here (DESIGN.md "Divergences") */
    int64_t e = r->stop < (uint64_t)g_arr_len ? r->stop : g_arr_len;
    for (int64_t p = r->start; p < e; p++) { int32_t v = g_arr[p]; for (int i = 0; i < nthr; i++) { if ((int64_t)v < thr[i]) break; counts[i]++; } }
    for (int i = 0; i < nthr; i++) fprintf(fh, "\t%" PRId64, counts[i]);
    fputc('\n', fh); free(counts);
}

static int64_t nim_toInt(double f) { return f >= 0 ? (int64_t)(f + 0.5) : (int64_t)(f - 0.5); }  /* system.toInt */

static FILE *xopen(const char *prefix, const char *suffix) { char p[4096]; snprintf(p, sizeof p, "%s%s", prefix, suffix); FILE *f = fopen(p, "w"); if (!f) die(1, "[mosdepth] could not open file:%s", p); static char *bufs[8]; static int nb; if (nb < 8) { bufs[nb] = (char *)xmalloc(1 << 20); setvbuf(f, bufs[nb++], _IOFBF, 1 << 20); } return f; }

int main(int argc, char **argv) {
    const char *chrom_arg = NULL, *by = NULL, *quant_arg = NULL, *thr_arg = NULL, *rg_arg = NULL;
    int threads = 0, no_per_base = 0, fast_mode = 0, fragment_mode = 0, use_median = 0, mapq = 0; int64_t min_len = -1, max_len = -1;
    uint16_t eflag = 1796, iflag = 0; const char *pos[2]; int npos = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
#define OPT(s, l) (!strcmp(a, s) || !strcmp(a, l))
#define VAL() (i + 1 < argc ? argv[++i] : (die(1, "error parsing arguments"), (char *)0))
        if (OPT("-t", "--threads")) threads = (int)nim_parse_int(VAL());
        else if (OPT("-c", "--chrom")) chrom_arg = VAL();
        else if (OPT("-b", "--by")) by = VAL();
        else if (OPT("-n", "--no-per-base")) no_per_base = 1;
        else if (OPT("-F", "--flag")) eflag = (uint16_t)nim_parse_int(VAL());
        else if (OPT("-i", "--include-flag")) iflag = (uint16_t)nim_parse_int(VAL());
        else if (OPT("-x", "--fast-mode")) fast_mode = 1;
        else if (OPT("-a", "--fragment-mode")) fragment_mode = 1;
        else if (OPT("-q", "--quantize")) quant_arg = VAL();
        else if (OPT("-Q", "--mapq")) mapq = (int)nim_parse_int(VAL());
        else if (OPT("-l", "--min-frag-len")) min_len = nim_parse_int(VAL());
        else if (OPT("-u", "--max-frag-len")) max_len = nim_parse_int(VAL());
        else if (OPT("-T", "--thresholds")) thr_arg = VAL();
        else if (OPT("-m", "--use-median")) use_median = 1;
        else if (OPT("-R", "--read-groups")) rg_arg = VAL();
        else if (a[0] == '-' && a[1]) die(1, "error parsing arguments");
        else if (npos < 2) pos[npos++] = a; else die(1, "error parsing arguments");
    }
    if (npos != 2) die(1, "error parsing arguments");
    const char *prefix = pos[0], *bam_path = pos[1];
    { const char *e = getenv("MOSDEPTH_PRECISION"); char *end; if (e && *e) { long v = strtol(e, &end, 10); g_precision = *end ? 2 : (int)v; } }   /* :28-32 */
    if (max_len < 0) max_len = INT64_MAX;                                                               /* :929-930 */
    if (max_len < min_len) { fprintf(stderr, "[mosdepth] error --max-frag-len was lower than --min-frag-len.\n"); return 2; }
    /* threshold_args :850-854 */
    int64_t *thr = NULL; int nthr = 0;
    if (thr_arg) { char *s = strdup(thr_arg), *tok = s; for (;;) { char *c = strchr(tok, ','); if (c) *c = 0; thr = (int64_t *)xrealloc(thr, (size_t)(nthr + 1) * 8); thr[nthr++] = nim_parse_int(tok); if (!c) break; tok = c + 1; } qsort(thr, (size_t)nthr, 8, i64_cmp); free(s); }
    if (fragment_mode && fast_mode) { fprintf(stderr, "[mosdepth] error only one of --fast-mode and --fragment-mode can be specified\n"); return 2; }
    if (!by && nthr) { fprintf(stderr, "[mosdepth] error --thresholds can only be used when --by is specified.\n"); return 2; }
    size_t bl = strlen(bam_path);
    if (bl > 5 && !strcmp(bam_path + bl - 5, ".cram")) { fprintf(stderr, "[mosdepth] ERROR: specify a reference file (or set REF_PATH env var) for decoding CRAM\n"); return 1; } /* :857-862 */

    bam_t bam; bam_open(&bam, bam_path, threads);
    index_t ix; if (index_load(&ix, bam_path) != 0) { fprintf(stderr, "[mosdepth] error alignment file must be indexed\n"); return 2; }  /* :969-971 */

    /* region_line_to_region :211-227 -> only the chromosome name is used (:706, Q9) */
    char *chrom = NULL;
    if (chrom_arg && *chrom_arg && strcmp(chrom_arg, "nil")) {
        chrom = strdup(chrom_arg);
        if (strchr(chrom, ':')) { /* rsplit({':','-'}, maxsplit=2): last two separators split off stop and start */
            int seps = 0; for (char *p = chrom + strlen(chrom) - 1; p >= chrom; p--) if (*p == ':' || *p == '-') { if (++seps == 2) { *p = 0; break; } }
            if (seps < 2) { char *p = strrchr(chrom, ':'); if (p) *p = 0; }
        }
        int ok = 0; for (int i = 0; i < bam.n_targets; i++) if (!strcmp(bam.targets[i].name, chrom)) ok = 1;
        if (!ok) { fprintf(stderr, "[mosdepth] chromosome %s not found\n", chrom); return 1; }   /* check_chrom :842-848 */
    }
    int64_t *quant = NULL; int nq = get_quantize_args(quant_arg, &quant);
    cov_opts co; memset(&co, 0, sizeof co); co.mapq = mapq; co.min_len = min_len; co.max_len = max_len; co.eflag = eflag; co.iflag = iflag; co.fast_mode = fast_mode; co.fragment_mode = fragment_mode;
    if (rg_arg) { char *s = strdup(rg_arg), *tok = s; for (;;) { char *c = strchr(tok, ','); if (c) *c = 0; co.read_groups = (char **)xrealloc(co.read_groups, (size_t)(co.n_rg + 1) * sizeof(char *)); co.read_groups[co.n_rg++] = strdup(tok); if (!c) break; tok = c + 1; } free(s); }

    /* main() mosdepth.nim:589-840 */
    depth_stat chrom_region_stat, chrom_stat, global_region_stat, global_stat;
    ds_clear(&chrom_region_stat); ds_clear(&chrom_stat); ds_clear(&global_region_stat); ds_clear(&global_stat);
    dist_t region_distribution, global_distribution, chrom_region_distribution, chrom_global_distribution;
    dist_new(&region_distribution, 512); dist_new(&global_distribution, 512); dist_new(&chrom_region_distribution, 512); dist_new(&chrom_global_distribution, 512);
    FILE *fbase = NULL, *fquant = NULL, *fthr = NULL, *fregion = NULL, *fh_gd, *fh_sum, *fh_rd = NULL;
    if (!no_per_base) fbase = xopen(prefix, ".per-base.bed");
    if (nq) fquant = xopen(prefix, ".quantized.bed");
    if (nthr) { fthr = xopen(prefix, ".thresholds.bed"); fputs("#chrom\tstart\tend\tregion", fthr); for (int i = 0; i < nthr; i++) fprintf(fthr, "\t%" PRId64 "X", thr[i]); fputc('\n', fthr); } /* :559-563 */
    fh_gd = xopen(prefix, ".mosdepth.global.dist.txt"); fh_sum = xopen(prefix, ".mosdepth.summary.txt");
    uint32_t window = 0; bed_table *bed = NULL; int by_digit = 0;
    if (by) { fh_rd = xopen(prefix, ".mosdepth.region.dist.txt"); fregion = xopen(prefix, ".regions.bed");
        by_digit = 1; for (const char *p = by; *p; p++) if (*p < '0' || *p > '9') by_digit = 0;
        if (by_digit) window = (uint32_t)nim_parse_int(by); else bed = bed_to_table(by); }
    count_stat cs; cs.len = use_median ? 65536 : 0; cs.n = 0; cs.counts = (uint32_t *)xcalloc((size_t)(cs.len ? cs.len : 1), 4);
    char **lookup = nq ? make_lookup(quant, nq) : NULL;
    seen_t seen; seen_init(&seen);

    for (int ti = 0; ti < bam.n_targets; ti++) {
        target_t *target = &bam.targets[ti];
        if (chrom && strcmp(target->name, chrom)) continue;                      /* get_targets :482-488 */
        memset(chrom_global_distribution.d, 0, (size_t)chrom_global_distribution.len * 8);
        if (by) memset(chrom_region_distribution.d, 0, (size_t)chrom_region_distribution.len * 8);
        bed_chrom *bc = bed_find(bed, target->name);
        if (no_per_base && nthr == 0 && nq == 0 && bed != NULL && !bc) continue; /* :703-705 (Q8) */
        if (chrom && !bam.have_look && ti < ix.n_ref && ix.has[ti] && bam.bz.cur_blk >= 0) {
            /* -c: use the index like bam.query does, instead of streaming through earlier contigs */
            bgzf_seek(&bam.bz, ix.beg[ti]); bam.eof = 0;
        }
        int tid = coverage(&bam, ti, &co, &seen);
        if (tid != -2) to_coverage();                                            /* :712-713 */
        if (by) {
            double me = 0; uint32_t L = (uint32_t)g_arr_len;
            /* region_gen :390-397 */
            region_t wr; wr.name = (char *)""; uint32_t wstart = 0; int widx = 0, wdone = 0;
            for (;;) {
                const region_t *r;
                if (!bed) {  /* window_gen :382-388 (uint32 arithmetic) */
                    if (wdone) break;
                    if ((uint32_t)(wstart + window) < target->length) { wr.start = wstart; wr.stop = wstart + window; wstart += window; }
                    else { if (wstart == target->length) break; wr.start = wstart; wr.stop = target->length; wdone = 1; }
                    r = &wr;
                } else { if (!bc || widx >= bc->n) break; r = &bc->r[widx++]; }
                if (tid != -2) {
                    me = imean(r->start, r->stop, &cs);                          /* :722 */
                    uint32_t a = r->start < L ? r->start : L, b = L < r->stop ? L : r->stop;
                    chrom_region_stat = ds_add(chrom_region_stat, ds_new(g_arr + a, b > a ? (int64_t)(b - a) : 0));  /* :723-724 */
                }
                if (!r->name || !*r->name) fprintf(fregion, "%s\t%u\t%u\t%.*f\n", target->name, r->start, r->stop, g_precision, me);
                else fprintf(fregion, "%s\t%u\t%u\t%s\t%.*f\n", target->name, r->start, r->stop, r->name, g_precision, me);
                if (tid != -2) {
                    if (by_digit) { int64_t bin = nim_toInt(me), hi = chrom_region_distribution.len - 1; chrom_region_distribution.d[bin < hi ? bin : hi] += 1; } /* :737-739 */
                    else dist_inc(&chrom_region_distribution, r->start, r->stop);                           /* :741 */
                }
                write_thresholds(fthr, tid, target->name, thr, nthr, r);         /* :743 */
            }
            if (bc) bc->used = 1;                                                /* bed_regions.del :397 */
        }
        if (tid != -2) {                                                         /* :745-753 */
            dist_inc(&chrom_global_distribution, 0, (uint32_t)(g_arr_len - 1));
            chrom_stat = ds_new(g_arr, g_arr_len - 1);
            global_stat = ds_add(global_stat, chrom_stat);
            write_summary(target->name, chrom_stat, fh_sum);
            if (by) { char nm[4096]; snprintf(nm, sizeof nm, "%s_region", target->name); write_summary(nm, chrom_region_stat, fh_sum); }
            global_region_stat = ds_add(global_region_stat, chrom_region_stat);
            ds_clear(&chrom_region_stat);
        }
        write_distribution(target->name, &chrom_global_distribution, fh_gd);    /* :756-758 */
        if (by) write_distribution(target->name, &chrom_region_distribution, fh_rd);
        if (tid >= 0) { if (by) sum_into(&chrom_region_distribution, &region_distribution); sum_into(&chrom_global_distribution, &global_distribution); } /* :761-764 */
        if (!no_per_base) {                                                      /* :766-793 */
            if (tid == -2) fprintf(fbase, "%s\t0\t%u\t0\n", target->name, target->length);
            else write_per_base(fbase, target->name);
        }
        if (nq) {                                                                /* :794-805 */
            if (tid == -2 && quant[0] == 0) fprintf(fquant, "%s\t0\t%u\t%s\n", target->name, target->length, lookup[0]);
            else { if (tid == -2) continue; write_quantized(fquant, target->name, quant, nq, lookup); }
        }
    }
    write_summary("total", global_stat, fh_sum);                                 /* :807-813 */
    if (by) write_summary("total_region", global_region_stat, fh_sum);
    write_distribution("total", &global_distribution, fh_gd);
    if (by) { write_distribution("total", &region_distribution, fh_rd); fclose(fh_rd); }
    if (bed && !chrom) for (int i = 0; i < bed->n; i++) if (!bed->c[i].used) fprintf(stderr, "[mosdepth] warning chromosome:%s from bed with %d regions not found\n", bed->c[i].chrom, bed->c[i].n); /* :816-819 */
    if (fregion) fclose(fregion);
    if (fquant) fclose(fquant);
    if (fthr) fclose(fthr);
    if (fbase) fclose(fbase);
    fclose(fh_gd); fclose(fh_sum);
    if (getenv("MOSDEPTH_ORACLE_STATS")) fprintf(stderr, "[oracle] records=%" PRId64 "\n", g_records_total);
    bam.bz.stop = 1;
    return 0;
}
