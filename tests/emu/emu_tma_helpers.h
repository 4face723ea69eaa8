// emu_tma_helpers.h — TEST INFRASTRUCTURE (tests/emu): host stand-ins for the PTX helpers of sweep_tma.cuh (mbarrier, TMA loads / stores,
// bulk-group waits), spliced in by tests/emu/build_emu.py where the product source has its inline PTX. A TMA copy happens synchronously at
// the point of issue (a legal ordering of the asynchronous one: the kernel may not touch a stage before its mbarrier phase completes, nor
// reuse it before the bulk-group wait), with the tensor map's bounds handling: out-of-bounds elements of a load read as zero, out-of-bounds
// elements of a store are dropped. The emulated mbarrier counts transaction bytes like the device's, so an expect_tx that disagrees with
// the boxes actually loaded hangs here too (the scheduler reports it as a deadlock / the test times out).
#pragma once

namespace tbrm {

struct EmuBar {           // lives in the 8 bytes of the kernel's mbarrier word
    int32_t tx_pending;   // bytes still to arrive in the current phase
    uint16_t phase;       // completed phases
    uint8_t count;        // arrivals a phase needs
    uint8_t pending;      // arrivals still missing in the current phase
};
static_assert(sizeof(EmuBar) == 8, "an mbarrier is one 64-bit word");

inline void emu_bar_settle(EmuBar* b) {
    if (b->pending == 0 && b->tx_pending == 0) b->phase++, b->pending = b->count;
}
inline void mbar_init(uint64_t* bar, int count) {
    if (count < 1 || count > 255) tbrm_emu::fail("emulated mbarrier: arrival count out of range");
    EmuBar* b = (EmuBar*) bar;
    b->tx_pending = 0, b->phase = 0, b->count = b->pending = (uint8_t) count;
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {  // mbarrier.arrive.expect_tx: one arrival + the transaction bytes to wait for
    EmuBar* b = (EmuBar*) bar;
    if (b->pending == 0) tbrm_emu::fail("emulated mbarrier: more arrivals than the phase expects");
    b->tx_pending += (int32_t) bytes;
    b->pending--;
    emu_bar_settle(b);
}
inline void mbar_arrive(uint64_t* bar) {
    EmuBar* b = (EmuBar*) bar;
    if (b->pending == 0) tbrm_emu::fail("emulated mbarrier: more arrivals than the phase expects");
    b->pending--;
    emu_bar_settle(b);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    const volatile EmuBar* b = (const volatile EmuBar*) bar;
    while ((b->phase & 1u) == (parity & 1u)) tbrm_emu::spin_hint();
}
inline void mbar_wait_long(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
// bar.sync id, nthreads: the fibers of a block run on one OS thread and switch only at barriers / spin hints
inline void named_bar_sync(int id, int nthreads) {
    static thread_local int arrived[16];
    static thread_local unsigned gen[16];
    if (id < 1 || id > 15) tbrm_emu::fail("emulated named barrier: id out of range");
    const unsigned g = gen[id];
    if (++arrived[id] == nthreads) {
        arrived[id] = 0, gen[id] = g + 1;
        return;
    }
    while (*(volatile unsigned*) &gen[id] == g) tbrm_emu::spin_hint();
}
inline void prefetch_l1(const void*) {}

inline void emu_tma_copy(void* smem, const CUtensorMap* map, int c0, int c1, int c2, bool load) {
    const tbrm_emu::TensorMap& m = *(const tbrm_emu::TensorMap*) map;
    if (m.magic != tbrm_emu::kTensorMapMagic) tbrm_emu::fail("TMA copy through something that is not an encoded tensor map");
    if (((uintptr_t) smem & 127u) != 0) tbrm_emu::fail("TMA: the shared-memory side of a tiled copy must be 128-byte aligned");
    const int c[3] = {c0, c1, c2};
    unsigned char* s = (unsigned char*) smem;
    for (uint32_t z = 0; z < m.box[2]; ++z)
        for (uint32_t y = 0; y < m.box[1]; ++y)
            for (uint32_t x = 0; x < m.box[0]; ++x, s += m.elem) {
                const long long gx = (long long) c[0] + x, gy = (long long) c[1] + y, gz = (long long) c[2] + z;
                const bool in = gx >= 0 && gy >= 0 && gz >= 0 && (uint64_t) gx < m.dims[0] && (uint64_t) gy < m.dims[1] && (uint64_t) gz < m.dims[2];
                unsigned char* g = (unsigned char*) m.base + (uint64_t) gx * m.elem + (uint64_t) gy * m.strides[0] + (uint64_t) gz * m.strides[1];
                if (load) {
                    if (in)
                        memcpy(s, g, m.elem);
                    else
                        memset(s, 0, m.elem);
                } else if (in) {
                    memcpy(g, s, m.elem);
                }
            }
}
inline void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    emu_tma_copy(dst, map, c0, c1, c2, true);
    const tbrm_emu::TensorMap& m = *(const tbrm_emu::TensorMap*) map;
    EmuBar* b = (EmuBar*) bar;
    b->tx_pending -= (int32_t) (m.box[0] * m.box[1] * m.box[2] * m.elem);
    if (b->tx_pending < 0) tbrm_emu::fail("emulated mbarrier: more transaction bytes than expect_tx announced");
    emu_bar_settle(b);
}
inline void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if (((uintptr_t) dst & 15u) || ((uintptr_t) src & 15u) || (bytes & 15u)) tbrm_emu::fail("bulk copy: 16-byte alignment / size");
    memcpy(dst, src, bytes);
    EmuBar* b = (EmuBar*) bar;
    b->tx_pending -= (int32_t) bytes;
    if (b->tx_pending < 0) tbrm_emu::fail("emulated mbarrier: more transaction bytes than expect_tx announced");
    emu_bar_settle(b);
}
inline void cp_async_16(void* dst, const void* src) {
    if (((uintptr_t) dst & 15u) || ((uintptr_t) src & 15u)) tbrm_emu::fail("cp.async: 16-byte alignment");
    const unsigned long long* s = (const unsigned long long*) src;
    unsigned long long* d = (unsigned long long*) dst;
    d[0] = __atomic_load_n(s, __ATOMIC_ACQUIRE), d[1] = __atomic_load_n(s + 1, __ATOMIC_ACQUIRE);
}
inline void cp_async_commit() {}
inline void cp_async_wait_but_one() {}
inline void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) { emu_tma_copy((void*) src, map, c0, c1, c2, false); }
inline void tma_commit() {}
template <int N>
inline void tma_wait_read() {}
template <int N>
inline void tma_wait_all() {}
inline void fence_async_smem() {}

}  // namespace tbrm
