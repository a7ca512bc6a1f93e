// CPU emulation of K0 (phylocsf_b200/csrc/pcsf_k0.cuh): runs the kernel's own indexing functions thread by thread,
// a CTA at a time, with the barrier between the two stages modelled by finishing stage 1 for every thread first.
// Test infrastructure only (tests/test_k0_emulation.py builds it with g++); the product runs the CUDA kernel.
#include <cstring>
#include <vector>

#include "pcsf_k0.cuh"

extern "C" int k0_emulate(const uint8_t* nt, int64_t nt_bytes, const int64_t* aln_off, const int32_t* aln_len,
                          const int64_t* region_off, int64_t nalign, int frames, int n_leaves, int tile_pos, int grid_y,
                          uint8_t* codes) {
    using namespace pcsf::k0;
    if (frames > MAX_FRAMES || tile_pos % 16 || tile_pos < 16 || grid_y < 1) return 1;
    const int64_t nwords = (nt_bytes + 3) >> 2;
    std::vector<uint32_t> ntw(nwords + 1, 0xA5A5A5A5u);  // what lies past the buffer must never matter
    memcpy(ntw.data(), nt, nt_bytes);
    std::vector<uint32_t> smem((smem_bytes(tile_pos, n_leaves, frames) + 3) / 4);
    for (int64_t a = 0; a < nalign; a++)
        for (int y = 0; y < grid_y; y++) {
            const int len = aln_len[a];
            for (int64_t p0 = (int64_t)y * tile_pos; p0 + 3 <= len; p0 += (int64_t)grid_y * tile_pos) {
                std::fill(smem.begin(), smem.end(), 0xEEEEEEEEu);  // stale tile contents must never reach the output
                Seg segs[MAX_FRAMES];
                for (int f = 0; f < MAX_FRAMES; f++) {
                    segs[f] = empty_seg();
                    if (f < frames) {
                        const int64_t r = a * frames + f, c0 = region_off[r];
                        segs[f] = make_seg(f, len, c0, (int)(region_off[r + 1] - c0), (int)p0, tile_pos, n_leaves);
                        if (segs[f].sbase + segs[f].nbytes > (f + 1) * slot_bytes(tile_pos, n_leaves)) return 2;  // a run must fit its slot
                    }
                }
                for (int tid = 0; tid < THREADS; tid++)
                    stage1_thread(tid, THREADS, ntw.data(), nwords, aln_off[a], len, n_leaves, (int)p0, tile_pos, segs, frames,
                                  (uint8_t*)smem.data());
                for (int tid = 0; tid < THREADS; tid++) stage2_thread(tid, THREADS, segs, frames, (const uint8_t*)smem.data(), codes);
            }
        }
    return 0;
}

extern "C" uint32_t k0_decode4(uint32_t w) { return pcsf::k0::decode4(w); }
extern "C" uint32_t k0_codon4(uint32_t a, uint32_t b, uint32_t c) { return pcsf::k0::codon4(a, b, c); }
extern "C" uint32_t k0_revcomp4(uint32_t w) { return pcsf::k0::revcomp4(w); }
extern "C" int k0_choose_tile_pos(int max_len, int n_leaves, int frames) { return pcsf::k0::choose_tile_pos(max_len, n_leaves, frames); }
