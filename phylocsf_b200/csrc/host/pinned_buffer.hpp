// pinned_buffer.hpp — page-locked staging memory for the command line's nucleotide buffers.
// The reader threads parse alignments straight into the buffers that pcsf_batch_upload_alignments_parts copies to
// the device. From pageable memory those copies go through the driver's bounce buffer - a CPU memcpy on the scoring
// thread at a few GB/s, with the device idle meanwhile (measured on the GPU box: the host pipeline alone parses 620 k
// alignments/s on 16 cores, but with scoring switched on the run spent its time in "batch hand-off"). Page-locked
// buffers are copied by DMA at PCIe speed while the host goes on.
// cudaHostAlloc is expensive (a driver call that pins pages), so memory is taken in 32 MB slabs through the C ABI
// (pcsf_host_alloc) and handed out in power-of-two size classes with free lists; buffers are recycled by the caller's
// pool, so after the first few batches nothing is allocated any more. Without a CUDA device (--strategy=nop) the slabs
// are ordinary memory.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../../include/phylocsf_b200.h"

namespace pcsf {
namespace host {

class PinnedArena {
  public:
    static PinnedArena& instance() {
        static PinnedArena a;
        return a;
    }
    void use_pinned(bool on) {
        std::lock_guard<std::mutex> lk(mu);
        pinned = on;
    }
    static int size_class(size_t bytes) {  // smallest k with (64 KB << k) >= bytes
        int k = 0;
        while (((size_t)65536 << k) < bytes) k++;
        return k;
    }
    // a block of at least `bytes`; cap receives its real size
    uint8_t* get(size_t bytes, size_t& cap) {
        const int k = size_class(bytes);
        cap = (size_t)65536 << k;
        std::lock_guard<std::mutex> lk(mu);
        if ((size_t)k < free_.size() && !free_[k].empty()) {
            uint8_t* p = free_[k].back();
            free_[k].pop_back();
            return p;
        }
        if (cap > slab_left) {  // the rest of the old slab stays unused (at most one block's worth per slab)
            const size_t want = cap > kSlab ? cap : kSlab;
            slab = (uint8_t*)raw_alloc(want);
            slab_left = slab ? want : 0;
            if (!slab) return nullptr;
        }
        uint8_t* p = slab;
        slab += cap;
        slab_left -= cap;
        return p;
    }
    void put(uint8_t* p, size_t cap) {
        if (!p) return;
        const int k = size_class(cap);
        std::lock_guard<std::mutex> lk(mu);
        if (free_.size() <= (size_t)k) free_.resize(k + 1);
        free_[k].push_back(p);
    }

  private:
    static constexpr size_t kSlab = (size_t)32 << 20;
    std::mutex mu;
    bool pinned = false;
    std::vector<std::vector<uint8_t*>> free_;
    uint8_t* slab = nullptr;
    size_t slab_left = 0;
    void* raw_alloc(size_t bytes) {
        if (pinned) {
            void* p = pcsf_host_alloc(bytes);
            if (p) return p;
            pinned = false;  // no device, or the driver refused: ordinary memory from here on
        }
        return std::malloc(bytes);
    }
};

// The subset of std::vector<uint8_t> the readers use, on arena blocks. Growing keeps the contents; new bytes are not
// initialised (the readers write every byte they hand on).
class NtBuffer {
  public:
    NtBuffer() = default;
    NtBuffer(const NtBuffer&) = delete;
    NtBuffer& operator=(const NtBuffer&) = delete;
    ~NtBuffer() { PinnedArena::instance().put(p_, cap_); }
    uint8_t* data() { return p_; }
    const uint8_t* data() const { return p_; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    void clear() { n_ = 0; }
    void reserve(size_t n) {
        if (n <= cap_) return;
        size_t cap = 0;
        uint8_t* q = PinnedArena::instance().get(n, cap);
        if (!q) throw std::bad_alloc();
        if (n_) std::memcpy(q, p_, n_);
        PinnedArena::instance().put(p_, cap_);
        p_ = q;
        cap_ = cap;
    }
    void resize(size_t n) {
        reserve(n);
        n_ = n;
    }

  private:
    uint8_t* p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

}  // namespace host
}  // namespace pcsf
