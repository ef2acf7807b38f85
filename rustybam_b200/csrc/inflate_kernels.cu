// inflate_kernels.cu — BGZF inflate on the device (reference: src/myio.rs:41-64, where `.bgz` input goes
// through gzp::BgzfSyncReader and `.gz` through flate2's GzDecoder before PafRecord::new ever sees a line).
//
// BGZF (what bgzip writes) is a series of gzip members of at most 64 KiB of payload each, every one an independent raw
// DEFLATE stream (RFC 1951) with its compressed size in a 'BC' extra field and CRC-32 + ISIZE in its trailer.  The host
// hops from header to header (one table row per block), the compressed bytes cross PCIe as they are in the file — a third
// of the text for CIGAR-heavy PAF — and ONE THREAD PER BLOCK inflates here: a 12.6 GB PAF is ~190 000 blocks, ten times
// the 148 x 160 threads that are resident at once (one warp per CTA, five CTAs per SM: the tables below take the shared
// memory), so block-level parallelism alone fills the machine.  Per thread: a 64-bit bit buffer
// refilled by byte loads, canonical-Huffman decoding through per-thread first-level look-up tables in SHARED memory (one load
// per symbol; interleaved across the lanes so that a warp's 32 independent look-ups are almost conflict-free) with the
// bit-serial (count, symbol) tables in local memory behind them for the rare longer codes — the bit-serial loop alone, a chain
// of dependent local-memory loads, ran at 1.6 MB/s per thread; all tables are rebuilt for every dynamic block —, LZ77 copies
// out of the thread's own output, CRC-32 of the block (slice-by-4 tables in
// shared memory) against the trailer.  Control flow is one loop with one symbol per trip, so the lanes of a warp
// reconverge at its head; they diverge only between literal and match.
// A plain single-member .gz is one long dependent stream: it stays with the host's zlib (documented, not silent: the
// caller asks rb_is_bgzf first).
#include <cstdint>

#include "rb_kernels.cuh"

namespace rb {

constexpr int INF_THREADS = 32;   // one warp per CTA: its look-up tables take 40 KB of shared memory
constexpr int LUT_L_BITS = 9, LUT_D_BITS = 7;  // first-level tables: literal/length codes of <= 9 bits, distance codes of <= 7
constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, MAXCODES = MAXLCODES + MAXDCODES, FIXLCODES = 288;

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitIn {
    const uint8_t* p;
    const uint8_t* end;
    unsigned long long buf;
    uint32_t cnt;
    bool over;  // read past the end of the block's payload
};
__device__ __forceinline__ void refill(BitIn& b) {
    while (b.cnt <= 56u) {
        uint32_t byte = 0;
        if (b.p < b.end) byte = *b.p;
        else if (b.p >= b.end + 8) { b.over = true; }  // (a few zero bytes of slack: the last code may be peeked past the end)
        b.p++;
        b.buf |= (unsigned long long)byte << b.cnt;
        b.cnt += 8u;
    }
}
__device__ __forceinline__ uint32_t getbits(BitIn& b, uint32_t n) {  // n <= 16
    if (b.cnt < n) refill(b);
    const uint32_t v = (uint32_t)b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

struct Huff {
    uint16_t* count;   // count[len] = number of symbols of that code length (0 .. MAXBITS)
    uint16_t* symbol;  // symbols ordered by code
};
// canonical Huffman tables from code lengths (RFC 1951 3.2.2); returns < 0 for an over-subscribed set, > 0 incomplete, 0 complete
__device__ int build(Huff& h, const uint8_t* length, int n) {
    uint16_t offs[MAXBITS + 1];
    for (int len = 0; len <= MAXBITS; len++) h.count[len] = 0;
    for (int s = 0; s < n; s++) h.count[length[s]]++;
    if (h.count[0] == n) return 0;  // no codes: complete, but decoding will fail
    int left = 1;
    for (int len = 1; len <= MAXBITS; len++) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int len = 1; len < MAXBITS; len++) offs[len + 1] = offs[len] + h.count[len];
    for (int s = 0; s < n; s++)
        if (length[s] != 0) h.symbol[offs[length[s]]++] = (uint16_t)s;
    return left;
}
// one symbol: the code is read bit by bit, most significant first, against the first code / count of every length
__device__ __forceinline__ int decode(BitIn& b, const Huff& h) {
    if (b.cnt < (uint32_t)MAXBITS) refill(b);
    int code = 0, first = 0, index = 0;
    unsigned long long buf = b.buf;
#pragma unroll 1
    for (int len = 1; len <= MAXBITS; len++) {
        code |= (int)(buf & 1ull);
        buf >>= 1;
        const int count = h.count[len];
        if (code - count < first) {
            b.buf = buf;
            b.cnt -= (uint32_t)len;
            return h.symbol[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;  // ran out of codes
}

// first-level table of a code: entry = symbol | length << 9 for every LUT index whose low `length` bits spell the code
// (DEFLATE packs Huffman codes most significant bit first into a least-significant-bit-first stream: the index holds the
// code bit-reversed); 0 = the code is longer than the table is wide.  `lut` is this thread's column of the interleaved table.
__device__ void build_lut(const Huff& h, uint16_t* lut, int bits) {
    for (int i = 0; i < (1 << bits); i++) lut[i * INF_THREADS] = 0;
    int first = 0, index = 0;
    for (int len = 1; len <= bits; len++) {
        const int count = h.count[len];
        for (int k = 0; k < count; k++) {
            const uint32_t sym = h.symbol[index + k];
            const uint32_t rev = __brev((uint32_t)(first + k)) >> (32 - len);
            for (uint32_t j = rev; j < (1u << bits); j += (1u << len)) lut[j * INF_THREADS] = (uint16_t)(sym | ((uint32_t)len << 9));
        }
        index += count;
        first = (first + count) << 1;
    }
}
__device__ __forceinline__ int decode_fast(BitIn& b, const Huff& h, const uint16_t* lut, int bits) {
    if (b.cnt < (uint32_t)MAXBITS) refill(b);
    const uint32_t e = lut[((uint32_t)b.buf & ((1u << bits) - 1u)) * INF_THREADS];
    const uint32_t len = e >> 9;
    if (len) {
        b.buf >>= len;
        b.cnt -= len;
        return (int)(e & 511u);
    }
    return decode(b, h);
}

__global__ void __launch_bounds__(INF_THREADS)
k_inflate_bgzf(const uint8_t* __restrict__ comp, const BgzfBlock* __restrict__ blk, uint32_t n_blk, uint8_t* __restrict__ out,
               unsigned long long* err /* min (block << 8 | code), ~0 = none */) {
    __shared__ uint32_t s_crc[4][256];
    __shared__ uint16_t s_lut_l[(1 << LUT_L_BITS) * INF_THREADS], s_lut_d[(1 << LUT_D_BITS) * INF_THREADS];
    uint16_t* lut_l = s_lut_l + threadIdx.x;
    uint16_t* lut_d = s_lut_d + threadIdx.x;
    for (int i = threadIdx.x; i < 256; i += INF_THREADS) {  // slice-by-4 CRC-32 tables (reflected polynomial 0xEDB88320)
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        s_crc[0][i] = c;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += INF_THREADS) {
        uint32_t c = s_crc[0][i];
        for (int t = 1; t < 4; t++) {
            c = s_crc[0][c & 0xFFu] ^ (c >> 8);
            s_crc[t][i] = c;
        }
    }
    __syncthreads();
    const uint32_t bi = blockIdx.x * INF_THREADS + threadIdx.x;
    if (bi >= n_blk) return;
    const BgzfBlock B = blk[bi];
    uint8_t* dst = out + B.out_off;
    const uint32_t out_len = B.out_len;
    uint32_t o = 0;
    int fail = 0;
    BitIn in;
    in.p = comp + B.cdata; in.end = in.p + B.clen; in.buf = 0; in.cnt = 0; in.over = false;

    uint16_t lcount[MAXBITS + 1], lsym[FIXLCODES], dcount[MAXBITS + 1], dsym[MAXDCODES];
    uint8_t lengths[MAXCODES + 2];
    Huff lencode{lcount, lsym}, distcode{dcount, dsym};

    int last = 0;
    while (!last && !fail) {
        last = (int)getbits(in, 1);
        const uint32_t type = getbits(in, 2);
        if (type == 0) {  // stored: skip to the byte boundary, LEN, NLEN, bytes
            in.buf >>= (in.cnt & 7u);
            in.cnt &= ~7u;
            const uint32_t len = getbits(in, 16), nlen = getbits(in, 16);
            if ((len ^ 0xFFFFu) != nlen) { fail = 2; break; }
            if (o + len > out_len) { fail = 3; break; }
            for (uint32_t i = 0; i < len; i++) dst[o + i] = (uint8_t)getbits(in, 8);
            o += len;
            continue;
        }
        if (type == 3) { fail = 4; break; }
        if (type == 1) {  // fixed codes (RFC 1951 3.2.6)
            int s = 0;
            for (; s < 144; s++) lengths[s] = 8;
            for (; s < 256; s++) lengths[s] = 9;
            for (; s < 280; s++) lengths[s] = 7;
            for (; s < FIXLCODES; s++) lengths[s] = 8;
            build(lencode, lengths, FIXLCODES);
            for (s = 0; s < MAXDCODES; s++) lengths[s] = 5;
            build(distcode, lengths, MAXDCODES);
        } else {  // dynamic codes (3.2.7)
            const int nlen = (int)getbits(in, 5) + 257, ndist = (int)getbits(in, 5) + 1, ncode = (int)getbits(in, 4) + 4;
            if (nlen > MAXLCODES || ndist > MAXDCODES) { fail = 5; break; }
            int idx = 0;
            for (; idx < ncode; idx++) lengths[c_clen_order[idx]] = (uint8_t)getbits(in, 3);
            for (; idx < 19; idx++) lengths[c_clen_order[idx]] = 0;
            if (build(lencode, lengths, 19) != 0) { fail = 6; break; }  // the code-length code must be complete
            idx = 0;
            while (idx < nlen + ndist) {
                int sym = decode(in, lencode);
                if (sym < 0) { fail = 7; break; }
                if (sym < 16) lengths[idx++] = (uint8_t)sym;
                else {
                    int len = 0, rep;
                    if (sym == 16) {
                        if (idx == 0) { fail = 8; break; }
                        len = lengths[idx - 1];
                        rep = 3 + (int)getbits(in, 2);
                    } else if (sym == 17) rep = 3 + (int)getbits(in, 3);
                    else rep = 11 + (int)getbits(in, 7);
                    if (idx + rep > nlen + ndist) { fail = 9; break; }
                    while (rep--) lengths[idx++] = (uint8_t)len;
                }
            }
            if (fail) break;
            if (lengths[256] == 0) { fail = 10; break; }  // no end-of-block code
            // (distance lengths are copied out before the literal/length table is built over the front of `lengths`)
            uint8_t dl[MAXDCODES];
            for (int s = 0; s < ndist; s++) dl[s] = lengths[nlen + s];
            int e = build(lencode, lengths, nlen);
            if (e < 0 || (e > 0 && nlen - lcount[0] != 1)) { fail = 11; break; }  // incomplete only if a single code
            e = build(distcode, dl, ndist);
            if (e < 0 || (e > 0 && ndist - dcount[0] != 1)) { fail = 12; break; }
        }
        build_lut(lencode, lut_l, LUT_L_BITS);
        build_lut(distcode, lut_d, LUT_D_BITS);
        // ---- symbols of this block ----
        for (;;) {
            const int sym = decode_fast(in, lencode, lut_l, LUT_L_BITS);
            if (sym < 0) { fail = 13; break; }
            if (sym < 256) {
                if (o >= out_len) { fail = 3; break; }
                dst[o++] = (uint8_t)sym;
            } else if (sym == 256) {
                break;
            } else {
                const int ls = sym - 257;
                if (ls >= 29) { fail = 14; break; }
                const uint32_t len = c_len_base[ls] + getbits(in, c_len_extra[ls]);
                const int ds = decode_fast(in, distcode, lut_d, LUT_D_BITS);
                if (ds < 0 || ds >= 30) { fail = 15; break; }
                const uint32_t dist = c_dist_base[ds] + getbits(in, c_dist_extra[ds]);
                if (dist > o) { fail = 16; break; }          // reaches in front of the block: BGZF blocks are self-contained
                if (o + len > out_len) { fail = 3; break; }
                // LZ77 copy out of the thread's own output.  The bytes come back from L2 (stores are written through, a load
                // right behind them misses L1): a byte-by-byte loop pays that round trip per byte — it bounded the first version
                // of this kernel at 1.7 MB/s per thread — so up to 16 loads are issued before the first store.  A chunk never
                // reads a byte it is about to write: it is at most `dist` long, the pattern of an overlapping match (dist < len)
                // repeats from what earlier chunks stored... which the next chunk's loads see, program order within a thread.
                const uint8_t* src = dst + o - dist;
                uint8_t* q = dst + o;
                if (dist == 1u) {  // run of one byte (very common in zero-padded or digit-heavy text): one load
                    const uint8_t c = src[0];
                    for (uint32_t i = 0; i < len; i++) q[i] = c;
                } else {
                    for (uint32_t i = 0; i < len;) {
                        uint32_t chunk = len - i;
                        chunk = chunk > dist ? dist : chunk;
                        chunk = chunk > 16u ? 16u : chunk;
                        uint8_t t[16];
#pragma unroll
                        for (int k = 0; k < 16; k++)
                            if ((uint32_t)k < chunk) t[k] = src[i + k];
#pragma unroll
                        for (int k = 0; k < 16; k++)
                            if ((uint32_t)k < chunk) q[i + k] = t[k];
                        i += chunk;
                    }
                }
                o += len;
            }
            if (in.over) { fail = 17; break; }
        }
    }
    if (!fail && o != out_len) fail = 18;  // ISIZE of the trailer
    if (!fail) {  // CRC-32 of the payload against the trailer
        uint32_t c = 0xFFFFFFFFu, i = 0;
        while (i < out_len && ((uintptr_t)(dst + i) & 3u)) { c = s_crc[0][(c ^ dst[i]) & 0xFFu] ^ (c >> 8); i++; }
        for (; i + 4 <= out_len; i += 4) {
            c ^= *reinterpret_cast<const uint32_t*>(dst + i);
            c = s_crc[3][c & 0xFFu] ^ s_crc[2][(c >> 8) & 0xFFu] ^ s_crc[1][(c >> 16) & 0xFFu] ^ s_crc[0][c >> 24];
        }
        for (; i < out_len; i++) c = s_crc[0][(c ^ dst[i]) & 0xFFu] ^ (c >> 8);
        if ((c ^ 0xFFFFFFFFu) != B.crc) fail = 19;
    }
    if (fail) atomicMin(err, ((unsigned long long)bi << 8) | (unsigned long long)fail);
}

void launch_inflate_bgzf(const uint8_t* comp, const BgzfBlock* blk, uint32_t n_blk, uint8_t* out, unsigned long long* err, cudaStream_t s) {
    if (n_blk == 0) return;
    k_inflate_bgzf<<<(n_blk + INF_THREADS - 1) / INF_THREADS, INF_THREADS, 0, s>>>(comp, blk, n_blk, out, err);
}

}  // namespace rb
