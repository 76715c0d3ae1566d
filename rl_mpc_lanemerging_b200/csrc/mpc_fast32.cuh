// fast32_kernel: the bounded first attempt of the fused gap-evaluation with 32-bit keys (round 2).  Included by mpc_fast.cu.
//
// Why: B200 has a native 32-bit shared-memory min (ATOMS.MIN, 2.3 cycles per warp instruction per SM -- the rate of a plain
// store; tools/microbench3.cu) but no 64-bit one: the 64-bit word of fast_pull_kernel is min-combined with pre-read + compare +
// ATOMS.CAS.64 + check + redo, 4x the cost and the source of the divergent lost-race paths (profiles/r01_fast_pull_h50.txt).
//
// Same DP, same winner at every cell, different encoding:
//   * labels are 32-bit fixed point (2^-f32_frac, DevParams) under a cost bound that keeps them below 0xff000000; the bounded
//     pass is exact for every problem whose answer lies below the bound (mpc_fast.cu, note "Cost bound"), the others -- plans
//     that cross a penalty zone or do not reach the horizon -- are handed to the 64-bit kernel through the fallback list;
//   * a successor is offered with ONE atom.shared.min.u32 on the key  [label >> 8 : 24 | 255 - v' : 8];
//   * the low 8 label bits do not fit the key.  They are carried exactly: the thread that finalises cell k of layer t (it
//     reads the winning key, so it knows v' and the predecessor cell k - v') rebuilds the full label from the predecessor's
//     state word -- [0xff | label & 255 | v | a + 128], written over the consumed key -- and the edge tables, adds the cell's
//     penalty, and leaves its own state word behind.  Keys (< 0xff000000) and state words (>= 0xff000000) share three arrays
//     that rotate with the layers; a stale state word reads as "empty" to the min, so nothing is ever cleared;
//   * two candidates whose labels agree in all but the low 8 bits are ordered by v' alone in the key.  The atomic returns the
//     previous key: when it lies in the same 24-bit bucket, the loser of the two goes to a small per-layer list and the
//     finalising thread decides between the key holder and the listed candidates on the exact labels (~5 of 2.4e5 offers per
//     problem at H=50).  Every candidate of the winning bucket is either the final holder or was listed when it lost to /
//     was displaced by a holder of the same bucket, so the winner is the exact minimum of (label, then larger v').
// The result equals the CPU model orc_solve_fast_model_q(f32_frac, 0, f32_bound) bit for bit (oracle/mpc_oracle.c).
#pragma once

#define F32_STATE 0xff000000u       // words >= this are state words / empty, words below are keys
#define F32_EMPTY 0xff000000u       // "no node": the state word with v = 0, a = -128.  No word of the arrays is ever >= 0xffffff00
                                    // (a state word has v <= 250), which is what lets a closed offer be issued with the key 0xffffffff
#define F32_EMPTY64 0xff000000ff000000ULL
#define F32_OVF 32                  // same-bucket candidates per layer (more: the problem is handed on)
#define F32_L2 512                  // cells of layers 1 and 2 (off-grid history) the prologue can hold

struct F32Tables { unsigned v[256], aj[8 + 32 * 16 + 8]; };     // aj padded: the state rebuild of layer 2 may look 8 entries outside
// Everything static in ONE struct: the hot loop addresses it as (one pinned base register) + (compile-time offset).  Left to itself
// nvcc rebuilds the shared-window address of every static array from S2UR / UMOV / ULEA at each use (~25 instructions per node).
struct __align__(16) F32Shared {
    FastShared FS;                  // first: staged with 16-byte vector stores
    F32Tables TB;
    unsigned l2full[F32_L2];        // full labels of layer 1, then of layer 2 (their edges are not tabulated)
    uint2 ovf[3][F32_OVF];          // layer t % 3: (cell, key) of candidates that lost to a key of their own bucket
    int ovfn[3];
    int chunk[3];
    unsigned long long best;
};

#ifdef MPC_HOST_EMU
__device__ __forceinline__ unsigned lds_u32(unsigned a) { emu::preempt_point(); emu::S().cell_reads++; return *emu::from_shared<unsigned>(a); }
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { emu::preempt_point(); *emu::from_shared<unsigned>(a) = v; }
__device__ __forceinline__ unsigned atoms_min_u32(unsigned a, unsigned val) {
    emu::preempt_point();
    emu::S().cas_issued++;
    unsigned *p = emu::from_shared<unsigned>(a), old = *p; if (val < old) *p = val; return old;
}
#else
__device__ __forceinline__ unsigned lds_u32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// (ptxas turns a predicated atom.shared with a result into a branch around it: closed offers are issued with a neutral key instead)
__device__ __forceinline__ unsigned atoms_min_u32(unsigned a, unsigned val) {
    unsigned old; asm volatile("atom.shared.min.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(val) : "memory"); return old;
}
#endif

// blocked cells of layer t from DENSE grids (K2): in an obstacle band, or closer than MIN_ALLOWED_DISTANCE to a distance-field edge
// (the penalty zone the bound excludes) -- the set LayerDesc::blk describes for the descriptor-fed kernel.  One warp per 32-cell
// word: coalesced reads of the mask bytes and the distances, one ballot.
template <class Prov>
__device__ __forceinline__ void build_blocked_bits_dense(const Prov &prov, int t, unsigned *bits, int klo, int khi, int num_s, double min_allowed,
                                                         int tid, int nth) {
    if (prov.vec_ok()) {
        // (uniform) aligned grids -- every grid mpc_build_grid writes: a warp takes 128 cells per trip, four per lane with one vector
        // load per array (two load instructions per thread instead of eight, a quarter of the trips), eight lanes OR their nibbles
        // into one word.  Words below klo >> 5 that the first group covers get their true bits too.
        const int w1 = khi >> 5, lane = tid & 31, nw = nth >> 5;
        for (int g = (klo >> 7) + (tid >> 5); g <= (khi >> 7); g += 2 * nw) {      // two groups per trip: independent loads in flight
            unsigned v[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int gg = g + u * nw;
                v[u] = gg <= (khi >> 7) ? prov.blocked4(t, (gg << 7) + (lane << 2), min_allowed, num_s) << ((lane & 7) << 2) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                unsigned x = v[u];
                x |= __shfl_xor_sync(0xffffffffu, x, 1); x |= __shfl_xor_sync(0xffffffffu, x, 2); x |= __shfl_xor_sync(0xffffffffu, x, 4);
                const int w = ((g + u * nw) << 2) + (lane >> 3);
                if ((lane & 7) == 0 && w <= w1) bits[w] = x;
            }
        }
        return;
    }
    constexpr int DEPTH = 4;                                                // words per warp and trip: that many independent loads in flight
    const int w1 = khi >> 5, lane = tid & 31, nw = nth >> 5;
    for (int w = (klo >> 5) + (tid >> 5); w <= w1; w += DEPTH * nw) {
        bool blocked[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; u++) {
            const int k = ((w + u * nw) << 5) + lane;
            blocked[u] = true;
            if (w + u * nw <= w1 && k < num_s) blocked[u] = prov.is_blocked_zone(t, k, min_allowed);
        }
#pragma unroll
        for (int u = 0; u < DEPTH; u++) {
            const unsigned m = __ballot_sync(0xffffffffu, blocked[u]);
            if (lane == 0 && w + u * nw <= w1) bits[w + u * nw] = m;
        }
    }
}

template <class Prov, bool WRAP, int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 192 ? 6 : MAXT <= 256 ? 4 : MAXT <= 384 ? 3 : MAXT <= 512 ? 2 : 1))
fast32_kernel(DevParams P, int B, SolveIO io, const LayerDesc *desc, const uint8_t *dense_ob, const void *dense_d, int dense_stride, int Wc) {
    constexpr bool DENSE = Prov::kDense;
#ifdef MPC_HOST_EMU
    unsigned char *const smem_raw = emu::S().dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    // dynamic shared memory: [F32Shared | clamp bits lo, hi | blocked bits 0, 1 | three arrays of Wc words]
    F32Shared &SH = *reinterpret_cast<F32Shared *>(smem_raw);
    FastShared &FS = SH.FS;
    F32Tables &TB = SH.TB;
    unsigned (&s_l2full)[F32_L2] = SH.l2full;
    uint2 (&s_ovf)[3][F32_OVF] = SH.ovf;
    int (&s_ovfn)[3] = SH.ovfn;
    int (&s_chunk)[3] = SH.chunk;
    unsigned long long &s_best = SH.best;
    BlockShared &S = FS.S;
    unsigned sbase = smem_u32(smem_raw);
#ifndef MPC_HOST_EMU
    asm volatile("" : "+r"(sbase));            // pinned: see F32Shared
#endif
    ClampBits CB;
    const int NW = (P.num_s_max + 31) >> 5;
    CB.lo = reinterpret_cast<unsigned *>(smem_raw + sizeof(F32Shared)); CB.hi = CB.lo + NW;
    unsigned *const blkbits[2] = {CB.hi + NW, CB.hi + 2 * NW + 2};      // blocked cells of layer L in blkbits[L & 1]
    const unsigned ab = sbase + (unsigned)sizeof(F32Shared) + 4u * (unsigned)(4 * NW + 4);    // three arrays of Wc 32-bit words: layer t lives in array t % 3
    uint16_t *bp = io.bp + (size_t)blockIdx.x * P.num_t * io.bp_stride;
    const int T = P.num_t, tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
    auto ring = [Wc](int k) -> int { return WRAP ? (k >= Wc ? k - Wc : k) : k; };
    for (int i = tid; i < 256; i += nth) TB.v[i] = P.vtab32[i];
    for (int i = tid; i < 8 + 512 + 8; i += nth) TB.aj[i] = (i >= 8 && i < 520) ? P.atab32[(i - 8) >> 4] + P.jtab32[(i - 8) & 15] : 0u;
    for (int k = tid; k < 3 * Wc; k += nth) sts_u32(ab + 4u * k, F32_EMPTY);
    if (io.B_dev) B = *io.B_dev;
    const unsigned bnd = P.f32_bound;
    const float kw = P.kw32;
    const unsigned tbv = sbase + (unsigned)offsetof(F32Shared, TB.v), tbaj = sbase + (unsigned)(offsetof(F32Shared, TB.aj) + 32), cba = sbase + (unsigned)sizeof(F32Shared);
    for (;;) {
        if (tid == 0) S.b = atomicAdd(io.work_counter, 1);
        __syncthreads();
        const int wi = S.b;
        if (wi >= B) break;
        const int b = io.subset ? io.subset[wi] : wi;
        if (b & (1 << 30)) {                  // (second shape) the bounded attempt has already failed: not for this kernel --
            if (tid == 0 && !io.skip_flagged) { const int q = atomicAdd(io.fallback_count, 1); io.fallback_list[q] = b; }   // passed on, unless the
            __syncthreads();                                                                                              // 64-bit kernel reads the same list
            continue;
        }
        SGrid g;
        double v0, a0;
        Prov prov;
        if constexpr (!DENSE) {
            g = make_sgrid(P, io.ego[4 * b], io.ego[4 * b + 1]); v0 = io.ego[4 * b + 2]; a0 = io.ego[4 * b + 3];
            prov.base = desc + (size_t)b * T; prov.sm = FS.layer;
        } else {
            g.s0 = io.s0[b]; g.ds = io.ds[b]; g.num_s = io.num_s[b]; v0 = io.v0[b]; a0 = io.a0[b];
            prov.ob_base = dense_ob + (size_t)b * T * dense_stride;
            prov.d_base = reinterpret_cast<decltype(prov.d_base)>(dense_d) + (size_t)b * T * dense_stride;
            prov.stride = dense_stride;
        }
        build_clamp_bits(P, g, CB);
        const double est_prev = __dsub_rn(g.s0, __dmul_rn(v0, P.p.t_disc));
        const double est_second = __dsub_rn(est_prev, __dmul_rn(P.p.t_disc, __dsub_rn(v0, __dmul_rn(a0, P.p.t_disc))));
        prov.load(1);
        if (tid == 0) {
            S.need_fallback = 0;
            for (int i = 0; i < 3; i++) { S.nlo[i] = INT_MAX; S.nhi[i] = -1; s_chunk[i] = 0; s_ovfn[i] = 0; }
            s_best = FX_EMPTY;
        }
        // ---- prologue: layers 0 -> 1 -> 2 have off-grid history (st_cy.pyx:329-330): exact fp64 edges, quantised; the few
        // layer-2 cells are min-combined as 64-bit words in a scratch row (array 0, not yet in use) and then turned into keys ----
        for (int k = tid; k < F32_L2; k += nth) { sts_u64(ab + 8u * k, FX_EMPTY); s_l2full[k] = 0xffffffffu; }
        int imin0, imax0;
        exact_window(P, g.s0, g.ds, g.s0, est_prev, est_second, imin0, imax0);
        if (imax0 > g.num_s) imax0 = g.num_s;
        prov.store(1);
        prov.load(2);
        __syncthreads();
        if (tid < imax0 - imin0) {                                // layer 1 (array 1): state word, v = k, a = 0
            const int kk = imin0 + tid;
            const double sn = g.sval(kk);
            bool ob; const double d = prov.eval_staged(1, kk, sn, ob);
            if (!ob && kk >= 256) S.need_fallback = 1;
            else if (!ob) {
                const unsigned long long l1 = (unsigned long long)__double2ll_rn(__dmul_rn(exact_cost(P, sn, g.s0, est_prev, est_second, d), P.f32_one));
                if (l1 <= bnd) {                                  // (a layer-1 cell inside a penalty zone is above the bound)
                    s_l2full[kk] = (unsigned)l1;
                    sts_u32(ab + 4u * (Wc + ring(kk)), F32_STATE | (((unsigned)l1 & 255u) << 16) | ((unsigned)kk << 8) | 128u);
                    bp[(size_t)1 * io.bp_stride + kk] = 0;
                    atomicMin(&S.nlo[1], kk); atomicMax(&S.nhi[1], kk);
                }
            }
        }
        prov.store(2);
        if (T > 3) prov.load(3);
        __syncthreads();
        if (T > 3) prov.store(3);
        if (T > 4) prov.load(4);
        bool fail = false;
        int dlo = 0, dhi = -1;
        if (S.nhi[1] < 0) fail = true;
        else {
            const int lo1 = S.nlo[1], hi1 = S.nhi[1], lme = P.lmax_exact;
            int mylo = INT_MAX, myhi = -1;
            for (int e = tid; e < (hi1 - lo1 + 1) * lme; e += nth) {      // one thread per (layer-1 node, successor)
                const int k1 = lo1 + e / lme, j = e % lme;
                const unsigned l1 = s_l2full[k1];
                if (l1 == 0xffffffffu) continue;
                const double s = g.sval(k1);
                int imin, imax;
                exact_window(P, g.s0, g.ds, s, g.s0, est_prev, imin, imax);
                const int kk = imin + j;
                if (kk >= imax || kk >= g.num_s) continue;
                bool blocked2;                                    // band or penalty zone (st_cy.pyx:383-384 / the bound)
                if constexpr (DENSE) blocked2 = prov.is_blocked_zone(2, kk, P.p.min_allowed_distance); else blocked2 = prov.is_blocked(2, kk);
                if (blocked2) continue;
                const int vn = kk - k1, an = vn - k1;
                if (vn > 255 || an < -16 || an > 15 || kk >= F32_L2) { S.need_fallback = 1; continue; }
                const unsigned long long tot = (unsigned long long)l1 + (unsigned long long)__double2ll_rn(__dmul_rn(exact_kin(P, g.sval(kk), s, g.s0, est_prev), P.f32_one));
                if (tot > bnd) continue;
                smem_min64(ab + 8u * kk, (tot << 16) | ((unsigned long long)(255 - vn) << 8) | (unsigned long long)(an + 128));
                mylo = min(mylo, kk); myhi = max(myhi, kk);
            }
            mylo = warp_min_i(mylo); myhi = warp_max_i(myhi);
            if (lane == 0 && myhi >= 0) { atomicMin(&S.nlo[2], mylo); atomicMax(&S.nhi[2], myhi); }
            __syncthreads();
            dlo = S.nlo[2]; dhi = S.nhi[2];
            if (dhi < 0) fail = true;
            for (int kk = tid; kk < F32_L2; kk += nth) {          // winners of layer 2 -> keys in array 2, full labels aside; scratch row back to "empty"
                const unsigned long long w2 = lds_u64(ab + 8u * kk);
                if (w2 != FX_EMPTY) {
                    const unsigned tot = (unsigned)(w2 >> 16);
                    s_l2full[kk] = tot;
                    sts_u32(ab + 4u * (2 * Wc + ring(kk)), (tot & ~255u) | (unsigned)((w2 >> 8) & 255u));
                }
                sts_u64(ab + 8u * kk, F32_EMPTY64);
            }
        }
        if (T > 4) prov.store(4);
        int bt = 0; unsigned long long best_word = 0ULL;
        if (!fail) {
            if (T > 5) prov.load(5);
            if (T > 3) {
                if constexpr (DENSE) {
                    if (tid == 0 && T > 4) prov.prefetch_span(4, dlo, min(dhi + 2 * P.vmax_c, g.num_s - 1));      // read by the build at t = 2
                    build_blocked_bits_dense(prov, 3, blkbits[1], dlo, min(dhi + P.vmax_c, g.num_s - 1), g.num_s, P.p.min_allowed_distance, tid, nth);
                }
                else build_blocked_bits(FS.layer[3], blkbits[1], dlo, min(dhi + P.vmax_c, g.num_s - 1), tid, nth);
            }
            // ---- layers 2 .. T-1: one barrier per layer; warps take 32-cell chunks of the layer's span ----
            for (int t = 2; t < T; t++) {
                const int s3 = t % 3, n3 = (t + 1) % 3, p3 = (t + 2) % 3;
                const unsigned cur = ab + 4u * (unsigned)(Wc * s3), nxt = ab + 4u * (unsigned)(Wc * n3), prv = ab + 4u * (unsigned)(Wc * p3);
                __syncthreads();                                  // offers into layer t, its span and list, staging of layer t+2, bits of t+1 are complete
                dlo = S.nlo[s3]; dhi = S.nhi[s3];
                if (dhi < 0) break;
                if (WRAP && dhi - dlo + 1 + (t == T - 1 ? 0 : P.vmax_c + 8) > Wc) { if (tid == 0) S.need_fallback = 1; break; }
                const int novf = min(s_ovfn[s3], F32_OVF);
                if (tid == 0) { S.nlo[p3] = INT_MAX; S.nhi[p3] = -1; s_chunk[n3] = 0; s_ovfn[p3] = 0; }      // what iteration t+1 accumulates into
                if (t + 3 < T) prov.store(t + 3);
                if (t + 4 < T) prov.load(t + 4);
                const bool last = (t == T - 1), first = (t == 2);
                const unsigned l2a = sbase + (unsigned)offsetof(F32Shared, l2full);
                if (t + 2 < T) {
                    if constexpr (DENSE) {
                        // next iteration's build (layer t + 3 over the span of layer t + 1 + 2 v_max): ask L2 for it now
                        if (tid == 0 && t + 3 < T) prov.prefetch_span(t + 3, dlo, min(dhi + 3 * P.vmax_c, g.num_s - 1));
                        build_blocked_bits_dense(prov, t + 2, blkbits[t & 1], dlo, min(dhi + 2 * P.vmax_c, g.num_s - 1), g.num_s, P.p.min_allowed_distance, tid, nth);
                    }
                    else build_blocked_bits(FS.layer[(t + 2) & 3], blkbits[t & 1], dlo, min(dhi + 2 * P.vmax_c, g.num_s - 1), tid, nth);
                }
                const unsigned edge0 = sbase + (unsigned)(offsetof(F32Shared, FS.layer) + offsetof(LayerSearch, edge)) + (unsigned)(t & 3) * (unsigned)sizeof(LayerSearch);
                const unsigned bucket0 = edge0 + (unsigned)(offsetof(LayerSearch, bucket_edge) - offsetof(LayerSearch, edge));
                const unsigned bwa = cba + 4u * (unsigned)(2 * NW + ((t + 1) & 1) * (NW + 2));
                uint16_t *bp_row = bp + (size_t)t * io.bp_stride;
#ifndef MPC_HOST_EMU
                asm volatile("" : "+l"(bp_row));
#endif
                unsigned long long mybest = FX_EMPTY;
                int mylo = INT_MAX, myhi = -1;
                const int nwarp = nth >> 5;
                int c = tid >> 5;
                for (;;) {
                    const int base = dlo + (c << 5);
                    if (base > dhi) break;
                    int cn = 0;
                    if (lane == 0) cn = nwarp + atomicAdd(&s_chunk[s3], 1);
                    const int k = base + lane, rk = ring(k);
                    unsigned w = F32_EMPTY;
                    if (k <= dhi) w = lds_u32(cur + 4u * rk);
                    double dcell = 0.0;                           // (dense grids) the cell's distance: in flight while the key is decoded
                    if constexpr (DENSE) { if (k <= dhi) dcell = prov.distance_at(t, k); }     // (one chunk ahead was measured slower)
                    if (w < F32_STATE) {
                        // the winning offer: v' from the key, the rest from the predecessor's state word
                        int v = 255 - (int)(w & 255u);
                        const unsigned stp = lds_u32(prv + 4u * ring(k - v));
                        int a = v - (int)((stp >> 8) & 255u);
                        const int jj = a - ((int)(stp & 255u) - 128);
                        unsigned low = ((stp >> 16) + lds_u32_nc(tbv + 4u * v) + lds_u32_nc(tbaj + 4u * ((a + 16) * 16 + (jj + 8)))) & 255u;
                        if (novf) {                               // candidates of the same bucket that lost on v': exact comparison
                            for (int i = 0; i < novf; i++) {
                                const uint2 oe = s_ovf[s3][i];
                                if (oe.x == (unsigned)k && ((oe.y ^ w) < 256u)) {
                                    const int vx = 255 - (int)(oe.y & 255u);
                                    const unsigned sx = lds_u32(prv + 4u * ring(k - vx));
                                    const int ax = vx - (int)((sx >> 8) & 255u), jx = ax - ((int)(sx & 255u) - 128);
                                    const unsigned lx = ((sx >> 16) + lds_u32_nc(tbv + 4u * vx) + lds_u32_nc(tbaj + 4u * ((ax + 16) * 16 + (jx + 8)))) & 255u;
                                    if (lx < low || (lx == low && vx > v)) { low = lx; v = vx; a = ax; }
                                }
                            }
                        }
                        unsigned full = (w & ~255u) | low;
                        if (first) full = lds_u32_nc(l2a + 4u * (unsigned)k);
                        unsigned label;
                        if constexpr (DENSE) {
                            // distance penalty: the cell's value in the dense grid (a finalised cell is outside the bands and zones)
                            label = full + fx_inv_penalty(kw, dcell);
                        } else {
                        // distance penalty: nearest distance-field edge on either side; edge[-1] / edge[n_edge] are -/+1e300
                        double sv = g.sval(k);
#ifndef MPC_HOST_EMU
                        asm volatile("" : "+d"(sv));              // (else it is computed twice: five fp64 instructions)
#endif
                        unsigned ea = edge0 + 8u * lds_u8_nc(bucket0 + (k >> MPC_BUCKET_SHIFT));
                        double hi = lds_f64_nc(ea);
                        while (hi < sv) { ea += 8u; hi = lds_f64_nc(ea); }
                        const double dl = __dsub_rn(sv, lds_f64_nc(ea - 8u)), dr = __dsub_rn(hi, sv);
                        label = full + fx_inv_penalty(kw, dr < dl ? dr : dl);
                        }
                        if (label <= bnd) {
                            MPC_EMU_COUNT_NODE();
                            sts_u32(cur + 4u * rk, F32_STATE | ((label & 255u) << 16) | ((unsigned)v << 8) | (unsigned)(a + 128));
                            bp_row[k] = (uint16_t)(k - v);
                            if (last) {
                                const unsigned long long key = ((unsigned long long)label << 16) | (unsigned long long)k;
                                mybest = key < mybest ? key : mybest;
                            } else {
                                int wlo, n;
                                int_window_sa(P, g.num_s, cba, NW, k, v, a, wlo, n);
                                if (n > 0) {
                                    const int vn = wlo - k, an = vn - v, jn = an - a;
                                    const unsigned wa = bwa + 4u * (unsigned)(wlo >> 5);
                                    const unsigned open = ~__funnelshift_r(lds_u32_nc(wa), lds_u32_nc(wa + 4u), wlo & 31) & ((1u << n) - 1u);
                                    mylo = min(mylo, wlo); myhi = max(myhi, wlo + n - 1);
                                    const unsigned tva = tbv + 4u * vn, taja = tbaj + 4u * ((an + 16) * 16 + (jn + 8));
                                    const unsigned tie = 255u - (unsigned)vn;
                                    const int r0 = ring(wlo);
                                    // offer e: label + V[v'] + A[a'] + J[j'] with v' = vn + e, a' = an + e, j' = jn + e
#define F32_KEY(E) (((label + lds_u32_nc(tva + 4u * (E)) + lds_u32_nc(taja + 68u * (E))) & ~255u) | (tie - (E)))
#define F32_LIST(CELL, OLD, KEY)            /* same bucket: list the loser (rare) */                                           \
                                    {                                                                                         \
                                        const int oi = atomicAdd(&s_ovfn[n3], 1);                                             \
                                        if (oi < F32_OVF) s_ovf[n3][oi] = make_uint2((unsigned)(CELL), (OLD) > (KEY) ? (OLD) : (KEY)); \
                                        else S.need_fallback = 1;                                                             \
                                    }
                                    if (!WRAP || r0 + 5 <= Wc) {
                                        // Straight line: five keys, five unconditional mins, one test.  A closed offer (blocked cell, or
                                        // past a window shorter than five cells) is issued with the key 0xffffffff: it changes nothing,
                                        // and no word in the arrays shares its upper 24 bits, so it cannot look like a same-bucket pair.
                                        const unsigned ra = nxt + 4u * (unsigned)r0;
                                        unsigned k0 = F32_KEY(0), k1 = F32_KEY(1), k2 = F32_KEY(2), k3 = F32_KEY(3), k4 = F32_KEY(4);
                                        k0 = (open & 1u) ? k0 : 0xffffffffu; k1 = (open & 2u) ? k1 : 0xffffffffu; k2 = (open & 4u) ? k2 : 0xffffffffu;
                                        k3 = (open & 8u) ? k3 : 0xffffffffu; k4 = (open & 16u) ? k4 : 0xffffffffu;
                                        const unsigned o0 = atoms_min_u32(ra, k0), o1 = atoms_min_u32(ra + 4u, k1), o2 = atoms_min_u32(ra + 8u, k2),
                                                       o3 = atoms_min_u32(ra + 12u, k3), o4 = atoms_min_u32(ra + 16u, k4);
                                        if (min(min(min(o0 ^ k0, o1 ^ k1), min(o2 ^ k2, o3 ^ k3)), o4 ^ k4) < 256u) {
                                            if ((o0 ^ k0) < 256u) F32_LIST(wlo, o0, k0)
                                            if ((o1 ^ k1) < 256u) F32_LIST(wlo + 1, o1, k1)
                                            if ((o2 ^ k2) < 256u) F32_LIST(wlo + 2, o2, k2)
                                            if ((o3 ^ k3) < 256u) F32_LIST(wlo + 3, o3, k3)
                                            if ((o4 ^ k4) < 256u) F32_LIST(wlo + 4, o4, k4)
                                        }
                                        for (int e = 5; e < n; e++)                 // (windows longer than 5 cells: other Settings)
                                            if ((open >> e) & 1u) {
                                                const unsigned key = F32_KEY(e), old = atoms_min_u32(nxt + 4u * (unsigned)ring(wlo + e), key);
                                                if ((old ^ key) < 256u) F32_LIST(wlo + e, old, key)
                                            }
                                    } else {                                        // the window crosses the end of the ring
                                        for (int e = 0; e < n; e++)
                                            if ((open >> e) & 1u) {
                                                const unsigned key = F32_KEY(e), old = atoms_min_u32(nxt + 4u * (unsigned)(r0 + e >= Wc ? r0 + e - Wc : r0 + e), key);
                                                if ((old ^ key) < 256u) F32_LIST(wlo + e, old, key)
                                            }
                                    }
#undef F32_KEY
#undef F32_LIST
                                }
                            }
                        } else sts_u32(cur + 4u * rk, F32_EMPTY);   // dropped by the bound: the key must not be seen again
                    }
                    c = __shfl_sync(FULL, cn, 0);
                }
                mylo = __reduce_min_sync(FULL, mylo); myhi = __reduce_max_sync(FULL, myhi);
                if (last) for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, mybest, o); mybest = x < mybest ? x : mybest; }
                if (lane == 0) {
                    if (mybest != FX_EMPTY) atomicMin(&s_best, mybest);
                    if (myhi >= 0) { atomicMin(&S.nlo[n3], mylo); atomicMax(&S.nhi[n3], myhi); }
                }
                if (last) {
                    __syncthreads();
                    if (s_best != FX_EMPTY) { bt = t; best_word = s_best; }
                }
            }
        }
        __syncthreads();
        if (bt < T - 1 || S.need_fallback) {
            // not finished here: offers that were never finalised may be left behind -> empty the arrays, hand the problem on.
            // Bit 30 of the list entry: the bounded attempt itself failed (the plan crosses a penalty zone or there is none),
            // the 64-bit kernel goes straight to its unbounded pass; without it (ring / table overflow) it starts over.
            const bool overflow = S.need_fallback != 0;
            for (int k = tid; k < 3 * Wc; k += nth) sts_u32(ab + 4u * k, F32_EMPTY);
            if (tid == 0) {
                const int q = atomicAdd(io.fallback_count, 1); io.fallback_list[q] = b | (overflow ? 0 : (1 << 30));
                if (overflow && io.overflow_count) atomicAdd(io.overflow_count, 1);
                if (!overflow && io.flagged_count) atomicAdd(io.flagged_count, 1);
            }
            __syncthreads();
            continue;
        }
        finish_problem(P, io, prov, &S, b, g, bt, (int)(best_word & 0xffff), (double)(best_word >> 16) / P.f32_one, bp, true);
    }
}

// bytes of dynamic shared memory in front of the three key / state arrays (the host adds 12 B per ring cell)
static size_t fast32_smem_head(int num_s_max) { return sizeof(F32Shared) + (size_t)4 * (4 * ((num_s_max + 31) >> 5) + 4); }

template <class Prov>
static cudaError_t launch_fast32_impl(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, const uint8_t *ob,
                                      const void *dist, int stride, cudaStream_t st) {
    cudaError_t e;
#define MPC_LAUNCH_F32(WRAPV, MAXTV)                                                                       \
    do {                                                                                                   \
        auto k = fast32_kernel<Prov, WRAPV, MAXTV>;                                                        \
        if ((e = set_smem(k, L.smem)) != cudaSuccess) return e;                                            \
        MPC_LAUNCH(k, L.grid, L.threads, L.smem, st, P, L.B, io, desc, ob, dist, stride, L.W);             \
    } while (0)
    if (L.threads <= 192) { if (L.wrap) MPC_LAUNCH_F32(true, 192); else MPC_LAUNCH_F32(false, 192); }
    else if (L.threads <= 256) { if (L.wrap) MPC_LAUNCH_F32(true, 256); else MPC_LAUNCH_F32(false, 256); }
    else if (L.threads <= 384) { if (L.wrap) MPC_LAUNCH_F32(true, 384); else MPC_LAUNCH_F32(false, 384); }
    else if (L.threads <= 512) { if (L.wrap) MPC_LAUNCH_F32(true, 512); else MPC_LAUNCH_F32(false, 512); }
    else { if (L.wrap) MPC_LAUNCH_F32(true, 1024); else MPC_LAUNCH_F32(false, 1024); }
#undef MPC_LAUNCH_F32
    return cudaGetLastError();
}

static cudaError_t launch_fast32_desc_impl(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const LayerDesc *desc, cudaStream_t st) {
    return launch_fast32_impl<FastDescProv>(P, L, io, desc, nullptr, nullptr, 0, st);
}

// K2: the same kernel on dense grids resident in HBM (mpc_solve_dense); fp32 or fp64 distances
static cudaError_t launch_fast32_dense_impl(const DevParams &P, const SolveLaunch &L, const SolveIO &io, const uint8_t *ob, const void *dist,
                                            int dist_f32, int stride, cudaStream_t st) {
    return dist_f32 ? launch_fast32_impl<FastDenseProv<float>>(P, L, io, nullptr, ob, dist, stride, st)
                    : launch_fast32_impl<FastDenseProv<double>>(P, L, io, nullptr, ob, dist, stride, st);
}
