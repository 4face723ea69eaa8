// simt_engine.cpp — TEST INFRASTRUCTURE: the fiber scheduler and the guarded device heap behind tests/emu/cuda_on_host.h (see there).
#include <sys/mman.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <thread>

#include "cuda_on_host.h"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace tbrm_emu {

[[noreturn]] void fail(const char* what) {
    fprintf(stderr, "tbrm_emu: %s (block %u,%u,%u thread %u,%u,%u)\n", what, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, threadIdx.y, threadIdx.z);
    abort();
}

int cooperative_supported() {
    const char* e = getenv("TBRM_EMU_COOPERATIVE");  // 0: behave like a device without cooperative launch (per-slice schedule everywhere)
    return (e && e[0] == '0') ? 0 : 1;
}

// ---- guarded device heap ---------------------------------------------------------------------------------------------------------------------
namespace {
std::mutex g_heap_mutex;
std::map<void*, std::pair<void*, size_t>> g_heap;  // user pointer -> mapping
}  // namespace

void* device_alloc(size_t bytes) {
    const size_t page = (size_t) sysconf(_SC_PAGESIZE);
    const size_t user = std::max<size_t>(16, (bytes + 15) / 16 * 16);
    const size_t body = (user + page - 1) / page * page;
    const size_t total = body + 2 * page;
    unsigned char* base = (unsigned char*) mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    mprotect(base, page, PROT_NONE);
    mprotect(base + page + body, page, PROT_NONE);
    unsigned char* p = base + page + body - user;  // the allocation ends at the guard page
    memset(p, 0xcd, user);                         // device memory is not zero-initialised: poison it
    std::lock_guard<std::mutex> lock(g_heap_mutex);
    g_heap[p] = {base, total};
    return p;
}

void device_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_heap_mutex);
    auto it = g_heap.find(p);
    if (it == g_heap.end()) fail("cudaFree of a pointer cudaMalloc did not return");
    munmap(it->second.first, it->second.second);
    g_heap.erase(it);
}

// ---- tensor maps ---------------------------------------------------------------------------------------------------------------------------
CUresult encode_tiled(CUtensorMap* out, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                      CUtensorMapFloatOOBfill) {
    const uint32_t elem = type == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1u : 4u;
    if (!out || rank != 3 || ((uintptr_t) base & 15u)) return CUDA_ERROR_INVALID_VALUE;
    for (int d = 0; d < 3; ++d)
        if (dims[d] == 0 || dims[d] > (1ull << 32) || box[d] == 0 || box[d] > 256 || elem_strides[d] != 1) return CUDA_ERROR_INVALID_VALUE;
    for (int d = 0; d < 2; ++d)
        if ((strides[d] & 15u) || strides[d] >= (1ull << 40)) return CUDA_ERROR_INVALID_VALUE;
    if ((box[0] * elem) & 15u) return CUDA_ERROR_INVALID_VALUE;
    TensorMap m;
    memset(&m, 0, sizeof(m));
    m.magic = kTensorMapMagic, m.base = base, m.elem = elem;
    for (int d = 0; d < 3; ++d) m.dims[d] = dims[d], m.box[d] = box[d];
    m.strides[0] = strides[0], m.strides[1] = strides[1];
    memset(out, 0, sizeof(*out));
    memcpy(out, &m, sizeof(m));
    return CUDA_SUCCESS;
}

// ---- fibers ----------------------------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kStackBytes = 128 << 10;  // per fiber; only the touched pages are ever backed

struct Warp {
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    unsigned long long slots[32] = {};
    bool lane_alive[32] = {};
};
struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    int linear = 0;
    bool done = false, started = false;
    long long wait_block = -1, wait_warp = -1;  // generation the fiber waits to pass (-1: runnable)
};
struct Block {
    uint3 bid;
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::vector<unsigned char> smem;
    const std::function<void()>* body = nullptr;
    ucontext_t sched;
    Fiber* cur = nullptr;
    std::vector<void*> stacks;
};
thread_local Block* t_block = nullptr;

void fiber_entry() {
    Block* b = t_block;
    Fiber* f = b->cur;
    (*b->body)();
    f->done = true;
    // thread exit: it no longer takes part in barriers
    Warp& w = b->warps[f->linear / 32];
    w.lane_alive[f->linear % 32] = false;
    if (--w.alive > 0 && w.arrived == w.alive) w.gen++, w.arrived = 0;
    if (--b->alive > 0 && b->arrived == b->alive) b->gen++, b->arrived = 0;
    swapcontext(&f->ctx, &b->sched);
}

void yield_to_scheduler() {
    Block* b = t_block;
    Fiber* f = b->cur;
    swapcontext(&f->ctx, &b->sched);
    // resumed: the scheduler restored threadIdx / blockIdx
}

void run_block(Block& b, dim3 block) {
    const int n = (int) (block.x * block.y * block.z);
    b.fibers.assign(n, Fiber());
    b.warps.assign((n + 31) / 32, Warp());
    b.alive = n, b.arrived = 0, b.gen = 0;
    while ((int) b.stacks.size() < n) {
        void* s = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
        if (s == MAP_FAILED) fail("out of memory for fiber stacks");
        b.stacks.push_back(s);
    }
    for (int i = 0; i < n; ++i) {
        Fiber& f = b.fibers[i];
        f.linear = i;
        f.tid = uint3{(unsigned) i % block.x, ((unsigned) i / block.x) % block.y, (unsigned) i / (block.x * block.y)};
        Warp& w = b.warps[i / 32];
        w.alive++, w.lane_alive[i % 32] = true;
    }
    t_block = &b;
    int remaining = n;
    while (remaining > 0) {
        bool progressed = false;
        for (int i = 0; i < n; ++i) {
            Fiber& f = b.fibers[i];
            if (f.done) continue;
            if (f.wait_block >= 0) {
                if ((unsigned) f.wait_block == b.gen) continue;
                f.wait_block = -1;
            }
            if (f.wait_warp >= 0) {
                if ((unsigned) f.wait_warp == b.warps[i / 32].gen) continue;
                f.wait_warp = -1;
            }
            if (!f.started) {
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = b.stacks[i];
                f.ctx.uc_stack.ss_size = kStackBytes;
                f.ctx.uc_link = nullptr;
                makecontext(&f.ctx, fiber_entry, 0);
                f.started = true;
            }
            b.cur = &f;
            threadIdx = f.tid, blockIdx = b.bid;
            swapcontext(&b.sched, &f.ctx);
            progressed = true;
            if (f.done) --remaining;
        }
        if (!progressed) fail("deadlock: every remaining thread of the block waits at a barrier that cannot complete");
    }
    t_block = nullptr;
}
}  // namespace

void sync_block() {
    Block* b = t_block;
    if (!b) fail("__syncthreads outside a kernel");
    Fiber* f = b->cur;
    if (++b->arrived == b->alive) {
        b->gen++, b->arrived = 0;
        return;
    }
    f->wait_block = b->gen;
    yield_to_scheduler();
}

void sync_warp() {
    Block* b = t_block;
    if (!b) fail("warp barrier outside a kernel");
    Fiber* f = b->cur;
    Warp& w = b->warps[f->linear / 32];
    if (++w.arrived == w.alive) {
        w.gen++, w.arrived = 0;
        return;
    }
    f->wait_warp = w.gen;
    yield_to_scheduler();
}

unsigned long long* warp_slot(int lane) { return &t_block->warps[t_block->cur->linear / 32].slots[lane]; }
int lane_id() { return t_block->cur->linear % 32; }
bool lane_alive(int lane) { return t_block->warps[t_block->cur->linear / 32].lane_alive[lane]; }
unsigned char* dynamic_smem() { return (unsigned char*) (((uintptr_t) t_block->smem.data() + 127u) & ~(uintptr_t) 127u); }

void spin_hint() {
    if (!t_block) return;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    yield_to_scheduler();
}

unsigned long long globaltimer_ns() {
    return (unsigned long long) std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void run_grid(dim3 grid, dim3 block, size_t smem_bytes, bool cooperative, const std::function<void()>& thread_body) {
    if (t_block) fail("nested kernel launch");
    const size_t nblocks = (size_t) grid.x * grid.y * grid.z;
    if (nblocks == 0 || block.x * block.y * block.z == 0) return;  // an invalid configuration on the device; the sources never launch one
    const dim3 saved_bd = blockDim, saved_gd = gridDim;
    auto one = [&](Block& b, size_t linear) {
        blockDim = block, gridDim = grid;
        b.bid = uint3{(unsigned) (linear % grid.x), (unsigned) ((linear / grid.x) % grid.y), (unsigned) (linear / ((size_t) grid.x * grid.y))};
        b.body = &thread_body;
        b.smem.assign(smem_bytes + 128, 0xcd);
        run_block(b, block);
    };
    if (!cooperative) {
        static thread_local Block b;  // keeps the fiber stacks between launches
        for (size_t i = 0; i < nblocks; ++i) one(b, i);
    } else {
        // all blocks co-resident: one OS thread per block (static __shared__ variables are thread_local), spin loops yield
        std::vector<std::thread> threads;
        std::vector<Block*> blocks(nblocks);
        for (size_t i = 0; i < nblocks; ++i) {
            blocks[i] = new Block();
            threads.emplace_back([&, i] { one(*blocks[i], i); });
        }
        for (auto& t : threads) t.join();
        for (Block* b : blocks) {
            for (void* s : b->stacks) munmap(s, kStackBytes);
            delete b;
        }
    }
    blockDim = saved_bd, gridDim = saved_gd;
}

}  // namespace tbrm_emu

cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
    const char* e = getenv("TBRM_EMU_TMA");  // 0: behave like a driver without tensor maps (the TMA-staged sweep reports "not handled")
    const bool have = !strcmp(symbol, "cuTensorMapEncodeTiled") && !(e && e[0] == '0');
    *fn = have ? (void*) &tbrm_emu::encode_tiled : nullptr;
    if (q) *q = have ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
