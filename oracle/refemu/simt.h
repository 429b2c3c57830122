// simt.h -- lock-step SIMT emulation used to run the reference's OWN OpenCL kernel text on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/hmme_oracle.h).  Nothing here restates the reference's
// algorithm: the kernel body of /root/reference/cl/sad.cl (calcSAD_AMP, :141-367) is compiled
// as C++ with `int` / `unsigned int` locals turned into 256-lane vectors, `__local` / `__global`
// pointers into gather/scatter views and `if (...)` into a lane mask, so that EVERY STATEMENT is
// executed by all work-items before the next one starts (read-all, then write-all).  That is the
// only deterministic reading of a kernel whose dependent local-memory accesses have no barriers
// (SURVEY.md App. B1), and it is what a 256-wide SIMD machine would do.
//
// Out-of-range accesses (the kernel reads local arrays up to index 511 of 256 and writes
// tempSad[..847] of 593, App. B2) are not UB here: reads return a poison value, writes are
// dropped, both are counted so that tests can assert no FINAL result depends on them.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <vector>

namespace simt {

typedef int I32;
typedef unsigned U32;
constexpr I32 MAXW = 256;
constexpr uint32_t POISON = 0xBAD0BAD0u;

struct Ctx {
    I32 W = 0;                 // work-items in the group
    I32 lsize[3] = {1, 1, 1};
    uint8_t mask[8][MAXW];     // mask stack (depth 0 = all active)
    I32 depth = 0;
    long oobReads = 0, oobWrites = 0;
};
inline Ctx& ctx() { static thread_local Ctx c; return c; }
inline const uint8_t* active() { return ctx().mask[ctx().depth]; }

struct V {
    uint32_t v[MAXW];
    V() {}
    V(I32 s) { for (I32 l = 0; l < MAXW; ++l) v[l] = (uint32_t)s; }
    V(U32 s) { for (I32 l = 0; l < MAXW; ++l) v[l] = s; }
    V(long s) { for (I32 l = 0; l < MAXW; ++l) v[l] = (uint32_t)s; }
    V(unsigned long s) { for (I32 l = 0; l < MAXW; ++l) v[l] = (uint32_t)s; }
};

#define SIMT_BINOP(op)                                                                          \
    inline V operator op(const V& a, const V& b) { V r; const I32 W = ctx().W;                 \
        for (I32 l = 0; l < W; ++l) r.v[l] = a.v[l] op b.v[l]; return r; }
SIMT_BINOP(+) SIMT_BINOP(-) SIMT_BINOP(*)
#undef SIMT_BINOP
// comparisons: every compared quantity in calcSAD_AMP is a small non-negative index, so signed and
// unsigned compare agree; use signed.
#define SIMT_CMP(op)                                                                            \
    inline V operator op(const V& a, const V& b) { V r; const I32 W = ctx().W;                 \
        for (I32 l = 0; l < W; ++l) r.v[l] = ((int32_t)a.v[l] op (int32_t)b.v[l]) ? 1u : 0u; return r; }
SIMT_CMP(<) SIMT_CMP(<=) SIMT_CMP(>) SIMT_CMP(>=) SIMT_CMP(==) SIMT_CMP(!=)
#undef SIMT_CMP
inline V operator&&(const V& a, const V& b) { V r; const I32 W = ctx().W;
    for (I32 l = 0; l < W; ++l) r.v[l] = (a.v[l] && b.v[l]) ? 1u : 0u; return r; }
inline V& operator+=(V& a, const V& b) { const I32 W = ctx().W; const uint8_t* m = active();
    for (I32 l = 0; l < W; ++l) if (m[l]) a.v[l] += b.v[l]; return a; }

// A view of a buffer (local or global) of element type T, indexed by a vector of lane indices.
template <typename T> struct Ptr;
template <typename T> struct Ref {
    const Ptr<T>* p; V idx;
    operator V() const;                       // gather (active lanes)
    const Ref& operator=(const V& val) const; // scatter (active lanes, lane order)
    const Ref& operator=(const Ref& o) const { V t = (V)o; return *this = t; }
};
template <typename T> struct Ptr {
    T* data = nullptr; size_t n = 0;
    Ptr() {}
    Ptr(T* d, size_t n_) : data(d), n(n_) {}
    Ref<T> operator[](const V& i) const { return Ref<T>{this, i}; }
};
template <typename T> Ref<T>::operator V() const {
    V r; const I32 W = ctx().W; const uint8_t* m = active();
    for (I32 l = 0; l < W; ++l) {
        if (!m[l]) { r.v[l] = POISON; continue; }
        const int32_t i = (int32_t)idx.v[l];
        if (i < 0 || (size_t)i >= p->n) { r.v[l] = POISON; ++ctx().oobReads; }
        else r.v[l] = (uint32_t)(int32_t)p->data[i];   // short sign-extends, unsigned passes through
    }
    return r;
}
template <typename T> const Ref<T>& Ref<T>::operator=(const V& val) const {
    const I32 W = ctx().W; const uint8_t* m = active();
    for (I32 l = 0; l < W; ++l) {
        if (!m[l]) continue;
        const int32_t i = (int32_t)idx.v[l];
        if (i < 0 || (size_t)i >= p->n) { ++ctx().oobWrites; continue; }
        p->data[i] = (T)val.v[l];
    }
    return *this;
}

// OpenCL built-in abs_diff(short, short) -> ushort, exact.
inline V abs_diff(const V& a, const V& b) { V r; const I32 W = ctx().W;
    for (I32 l = 0; l < W; ++l) { const int32_t x = (int16_t)a.v[l], y = (int16_t)b.v[l];
        r.v[l] = (uint32_t)(uint16_t)(x > y ? x - y : y - x); }
    return r; }

struct MaskGuard {
    bool first = true;
    explicit MaskGuard(const V& c) {
        Ctx& k = ctx(); const uint8_t* cur = k.mask[k.depth]; uint8_t* nxt = k.mask[k.depth + 1];
        for (I32 l = 0; l < k.W; ++l) nxt[l] = cur[l] && c.v[l];
        ++k.depth;
    }
    ~MaskGuard() { --ctx().depth; }
    bool once() { const bool f = first; first = false; return f; }
};

inline V local_id(I32 dim) { V r; Ctx& k = ctx();
    for (I32 l = 0; l < k.W; ++l) {
        const I32 x = l % k.lsize[0], y = (l / k.lsize[0]) % k.lsize[1], z = l / (k.lsize[0] * k.lsize[1]);
        r.v[l] = (uint32_t)(dim == 0 ? x : dim == 1 ? y : z);
    }
    return r; }

inline void begin_group(I32 lx, I32 ly) { Ctx& k = ctx(); k.lsize[0] = lx; k.lsize[1] = ly; k.lsize[2] = 1;
    k.W = lx * ly; k.depth = 0; std::memset(k.mask[0], 1, MAXW); }

}  // namespace simt
