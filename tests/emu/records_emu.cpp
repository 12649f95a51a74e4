// records_emu.cpp -- K2 (frame.cuh), K3/K5 and K4 (records.cuh) executed on the CPU under tests/emu/cuda_runtime.h, chunk by chunk in
// the launch order of msd_run (mosdepth_b200/csrc/msd_api.cu): BGZF blocks are inflated by zlib here (K1 has its own host replay),
// laid out as [carry | blocks] like the library's chunk buffers, framed, filtered, turned into +1/-1 events (default mode: with the
// mate join and its live list carried from chunk to chunk), and the unfinished tail is carried into the next chunk's buffer.  Then a
// plain prefix sum and run-length pass print per-base.bed, which tests/test_kernels_emu.py compares with the oracle's, byte for byte.
// Small chunks (a few blocks) make records and mates straddle chunk boundaries.  TEST INFRASTRUCTURE ONLY.
//   usage: records_emu <in.bam> <out.per-base.bed> <mode 0|1|2> <blocks_per_chunk> <max_ctas>
#include <cstdlib>
#include <string>
#include <zlib.h>
#include "../../include/mosdepth_b200.h"
#include "../../mosdepth_b200/csrc/records.cuh"

using namespace msd;
typedef unsigned long long ull;

struct Blk { std::vector<uint8_t> data; };

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: records_emu in.bam out.bed mode blocks_per_chunk max_ctas\n"); return 2; }
    const int mode = atoi(argv[3]); const size_t per_chunk = (size_t)atoi(argv[4]); const int max_ctas = atoi(argv[5]);
    // ---- host: read + inflate every BGZF block (stands in for K1), parse the BAM header -----------------------------------------
    std::vector<uint8_t> F; { FILE* f = fopen(argv[1], "rb"); if (!f) return 2; fseek(f, 0, SEEK_END); F.resize((size_t)ftell(f)); fseek(f, 0, SEEK_SET); if (fread(F.data(), 1, F.size(), f) != F.size()) return 2; fclose(f); }
    std::vector<Blk> blocks;
    for (size_t pos = 0; pos + 18 <= F.size();) {
        const uint32_t xlen = F[pos + 10] | (F[pos + 11] << 8), bsize = (F[pos + 16] | (F[pos + 17] << 8)) + 1u;
        uint32_t isz; memcpy(&isz, F.data() + pos + bsize - 4, 4);
        Blk b; b.data.resize(isz);
        if (isz) { z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15); zs.next_in = F.data() + pos + 12 + xlen; zs.avail_in = bsize - xlen - 20; zs.next_out = b.data.data(); zs.avail_out = isz;
                   const int r = inflate(&zs, Z_FINISH); inflateEnd(&zs); if (r != Z_STREAM_END) { fprintf(stderr, "zlib failed\n"); return 2; } blocks.push_back(std::move(b)); }
        pos += bsize;
    }
    std::vector<uint8_t> all; for (auto& b : blocks) all.insert(all.end(), b.data.begin(), b.data.end());
    if (all.size() < 12 || memcmp(all.data(), "BAM\1", 4)) return 2;
    int32_t l_text, n_ref; memcpy(&l_text, all.data() + 4, 4); size_t off = 8 + (size_t)l_text; memcpy(&n_ref, all.data() + off, 4); off += 4;
    std::vector<std::string> names; std::vector<uint32_t> tlen;
    for (int i = 0; i < n_ref; i++) { int32_t ln; memcpy(&ln, all.data() + off, 4); names.emplace_back((const char*)all.data() + off + 4, (size_t)ln - 1); int32_t l; memcpy(&l, all.data() + off + 4 + ln, 4); tlen.push_back((uint32_t)l); off += 8 + (size_t)ln; }
    size_t b0 = 0, acc0 = 0; while (b0 < blocks.size() && acc0 + blocks[b0].data.size() <= off) acc0 += blocks[b0++].data.size();     // block of the first record
    const uint32_t first_uoff = (uint32_t)(off - acc0);
    // ---- "device" state, sized like msd_api.cu sizes it (small) --------------------------------------------------------------------
    std::vector<uint64_t> depth_off((size_t)n_ref); ull cells = 0; for (int t = 0; t < n_ref; t++) { depth_off[(size_t)t] = cells; cells += (((ull)tlen[(size_t)t] + 1) * 4 + 127) / 128 * 32; }
    std::vector<int32_t> arena(cells + 64, 0);
    std::vector<uint32_t> range_lo((size_t)n_ref, 0), range_hi((size_t)n_ref, 0xffffffffu); std::vector<ull> n_beyond((size_t)n_ref, 0), nprefilter((size_t)n_ref, 0);
    ContigTable ct{depth_off.data(), tlen.data(), range_lo.data(), range_hi.data(), nullptr, n_beyond.data(), nullptr};
    FilterParams fp; memset(&fp, 0, sizeof fp); fp.mapq = 0; fp.min_len = -1; fp.max_len = INT64_MAX; fp.eflag = 1796; fp.iflag = 0; fp.mode = mode; fp.n_rg = 0; fp.rg_names = nullptr; fp.n_targets = n_ref;
    const uint32_t carry_cap = 4u << 20; const size_t chunk_infl = per_chunk * 65536 + 65536;
    std::vector<uint8_t> infl[2]; for (auto& v : infl) v.assign((size_t)carry_cap + chunk_infl + 4096, 0);
    const uint32_t rec_cap = (uint32_t)((chunk_infl + carry_cap) / 36 + 2), nbmax = (uint32_t)per_chunk + 4;
    std::vector<uint32_t> rel(nbmax + 2), guess(nbmax + 4), seg(nbmax + 4), bstart(nbmax + 4), land(nbmax + 4), count(nbmax + 4), first(nbmax + 4), cnt(nbmax + 4), recbase(nbmax + 4), recoff(rec_cap);
    ChunkState cs; memset(&cs, 0, sizeof cs); RunCounters rc; memset(&rc, 0, sizeof rc);
    std::vector<uint8_t> cls(rec_cap); std::vector<uint64_t> qhash(rec_cap); std::vector<int32_t> endpos(rec_cap); std::vector<uint32_t> nxt(rec_cap), slot(rec_cap);
    MateChunk mc{cls.data(), qhash.data(), endpos.data(), nxt.data(), slot.data()};
    const uint32_t live_cap = std::max<uint32_t>(rec_cap * 4u, 1024u); uint32_t slots = 1024; while (slots < 2ull * (rec_cap + live_cap) && slots < (1u << 26)) slots <<= 1;
    std::vector<ull> mt_word(slots); std::vector<uint32_t> mt_head(slots), mt_live(slots);
    MateTable mt{mt_word.data(), mt_head.data(), mt_live.data(), slots - 1};
    const uint32_t arena_cap = std::max<uint32_t>(live_cap * 64u, 1u << 20);
    std::vector<LiveEntry> le[2]; std::vector<uint8_t> la[2]; uint32_t ln[2] = {0, 0}, lu[2] = {0, 0}; LiveList live[2];
    for (int k = 0; k < 2; k++) { le[k].resize(live_cap); la[k].resize(arena_cap); live[k] = LiveList{le[k].data(), la[k].data(), &ln[k], &lu[k], live_cap, arena_cap}; }
    int live_cur = 0;
    const uint32_t base = carry_cap;
    // ---- the chunk loop of msd_run ----------------------------------------------------------------------------------------------------
    size_t k = 0; bool first_of_span = true; ull n_chunks = 0, n_segments = 0;
    for (size_t b = b0; b < blocks.size();) {
        uint8_t* buf = infl[k].data(); uint32_t nb = 0, total = 0;
        while (b < blocks.size() && nb < per_chunk) { rel[nb] = total; memcpy(buf + base + total, blocks[b].data.data(), blocks[b].data.size()); total += (uint32_t)blocks[b].data.size(); nb++; b++; }
        rel[nb] = total;
        const int nblk = (int)nb + 1;
        emu::launch((nb + 255) / 256, 256, [&]() { frame_setup_kernel(&cs, rel.data(), (int)nb, base, first_of_span ? first_uoff : 0u, total, first_of_span ? 1 : 0, bstart.data()); });
        emu::launch((unsigned)(nblk * 32 + 255) / 256, 256, [&]() { frame_guess_kernel(buf, &cs, bstart.data(), nblk, n_ref, guess.data()); });
        emu::launch(1, 1024, [&]() { frame_segments_kernel(&cs, guess.data(), nblk, seg.data()); });
        n_segments += cs.n_seg;
        emu::launch((unsigned)(nblk + 127) / 128, 128, [&]() { frame_spec_kernel(buf, &cs, seg.data(), nblk, land.data(), count.data()); });
        emu::launch(1, 1024, [&]() { frame_link_kernel(buf, &cs, seg.data(), nblk, land.data(), count.data(), first.data(), cnt.data(), recbase.data()); });
        emu::launch((unsigned)(nblk + 127) / 128, 128, [&]() { frame_emit_kernel(buf, &cs, first.data(), cnt.data(), recbase.data(), nblk, recoff.data()); });
        if (cs.n_rec > rec_cap) { fprintf(stderr, "rec_cap\n"); return 3; }
        emu::launch((unsigned)max_ctas, 256, [&]() { records_kernel(buf, recoff.data(), &cs, fp, ct, arena.data(), nprefilter.data(), mc, &rc); });
        if (mode == MSD_MODE_DEFAULT) {
            LiveList& L = live[live_cur]; LiveList& Ln = live[live_cur ^ 1];
            memset(mt_word.data(), 0xff, 8ull * slots); memset(mt_head.data(), 0xff, 4ull * slots); memset(mt_live.data(), 0xff, 4ull * slots);
            *Ln.n = 0; *Ln.arena_used = 0;
            emu::launch((unsigned)std::max(1, max_ctas / 4), 256, [&]() { mate_reinsert_kernel(mt, L, buf, recoff.data(), &rc); });
            emu::launch((unsigned)max_ctas, 256, [&]() { mate_link_kernel(mt, L, buf, recoff.data(), &cs, mc, 0, &rc, ct); });
            emu::launch((unsigned)max_ctas, 256, [&]() { mate_link_kernel(mt, L, buf, recoff.data(), &cs, mc, 1, &rc, ct); });
            emu::launch((unsigned)max_ctas, 256, [&]() { mate_replay_kernel(mt, L, Ln, buf, recoff.data(), &cs, mc, ct, arena.data(), &rc); });
            emu::launch((unsigned)std::max(1, max_ctas / 4), 256, [&]() { mate_carry_kernel(L, Ln, &rc); });
            live_cur ^= 1;
        }
        emu::launch(4, 256, [&]() { carry_copy_kernel(buf, infl[k ^ 1].data(), &cs, base, carry_cap, &rc.error); });
        if (cs.error || rc.error) { fprintf(stderr, "kernel error: chain %u, records %u\n", cs.error, rc.error); return 4; }
        first_of_span = false; k ^= 1; n_chunks++;
    }
    // ---- to_coverage + gen_depths on the host (K6 / K8 have their own emulation tests) ----------------------------------------------------
    FILE* o = fopen(argv[2], "wb"); if (!o) return 2;
    for (int t = 0; t < n_ref; t++) {
        const uint32_t tl = tlen[(size_t)t];
        if (nprefilter[(size_t)t] == 0) { fprintf(o, "%s\t0\t%u\t0\n", names[(size_t)t].c_str(), tl); continue; }
        int32_t* a = arena.data() + depth_off[(size_t)t]; uint32_t s = 0;
        for (uint32_t i = 0; i <= tl; i++) { s += (uint32_t)a[i]; a[i] = (int32_t)s; }
        for (uint32_t st = 0, i = 1; i <= tl; i++) if (i == tl || a[i] != a[st]) { fprintf(o, "%s\t%u\t%u\t%d\n", names[(size_t)t].c_str(), st, i, a[st]); st = i; }
    }
    fclose(o);
    printf("ok %llu chunks, %llu segments, %llu records, %llu events\n", n_chunks, n_segments, rc.n_records, rc.n_events);
    return 0;
}
