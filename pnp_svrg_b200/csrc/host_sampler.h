// Host side of the minibatch sampler (see host_sampler.cpp); internal to libpnp_b200.so.
#pragma once
#include <vector>

namespace pnp_host {

inline unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// one pass of the network of pnp::feistel_perm (csmri.cuh); the permutation is this pass iterated while x >= n
unsigned feistel_pass(unsigned x, unsigned n, unsigned key, int hb);
// out[j] = perm(first + j) for j in [0, cnt): feistel_pass iterated until the value is below n
void feistel_block(unsigned first, unsigned* out, int cnt, unsigned n, unsigned key, int hb);
// out[i] = support[perm(i)] (or perm(i) when support is null) for i in [lo, hi)
void sample_range(int* out, int lo, int hi, unsigned n, unsigned key, int hb, const int* support);


// Look-ahead queue of draws for mb_source='host': `ahead` worker threads (plain CPU code, no CUDA calls) write draw
// number c = 0, 1, 2, ... into buffers[c % n_buffers]; draw c may start once c < consumed + ahead.  The consumer
// takes the draws in order: wait_next() blocks until draw number `consumed` is complete and returns its slot,
// release() hands it out (consumed += 1).  The buffer of draw c is overwritten by draw c + n_buffers, which becomes
// claimable at release number c + n_buffers - ahead + 1: whoever reads the buffers asynchronously (an H2D copy)
// must make sure that read has finished before that release (pnp_host_draws_stage does, with one event per slot).
class DrawQueue {
public:
    DrawQueue(int n, int count, unsigned seed, const int* support, int* const* buffers, int n_buffers, int ahead);
    ~DrawQueue();
    DrawQueue(const DrawQueue&) = delete;
    DrawQueue& operator=(const DrawQueue&) = delete;
    int wait_next();
    void release();
    long long consumed() const { return consumed_; }          // consumer thread only
    int n_buffers() const { return (int)bufs_.size(); }
    int ahead() const { return ahead_; }
    int count() const { return count_; }
    bool has_support() const { return support_ != nullptr; }
    int* buffer(int slot) const { return bufs_[slot]; }

private:
    struct Impl;
    void worker();
    int n_, count_, hb_, ahead_;
    unsigned seed_;
    const int* support_;
    std::vector<int*> bufs_;
    long long consumed_ = 0;
    Impl* impl_;
};

}  // namespace pnp_host
