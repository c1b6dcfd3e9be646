// Host side of the minibatch sampler (see host_sampler.cpp); internal to libpnp_b200.so.
#pragma once

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

}  // namespace pnp_host
