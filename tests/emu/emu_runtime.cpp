// tests/emu/emu_runtime.cpp — TEST INFRASTRUCTURE ONLY (see shim/cuda_runtime.h).
//
// Fiber-based SIMT engine: one CTA at a time, every CUDA thread a fiber with its
// own stack, cooperative round-robin scheduling on ONE OS thread.  A fiber runs
// until it reaches a rendezvous (warp collective, __syncthreads, mbarrier wait);
// a full scheduling pass without any state change is reported as a deadlock
// (mismatched collectives / barriers), which on a GPU would be a hang.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <map>
#include <vector>

extern "C" void gzpb_emu_switch(void **save_sp, void **load_sp);
asm(R"(
.text
.globl gzpb_emu_switch
.type gzpb_emu_switch,@function
gzpb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size gzpb_emu_switch,.-gzpb_emu_switch
)");

namespace gzpb_emu {

Self *g_self = nullptr;

namespace {
constexpr size_t kStackBytes = 512 << 10;

struct Fiber {
    Self self;
    void *sp = nullptr;
    uint8_t *stack = nullptr;
    bool done = false;
};

struct WarpState {
    unsigned exist = 0, exited = 0;
    unsigned mask = 0, arrived = 0, departed = 0;
    bool releasing = false;
    int op = 0;
    uint64_t val[32];
};

std::vector<Fiber> g_fibers;
std::vector<WarpState> g_warps;
void *g_sched_sp = nullptr;
const std::function<void()> *g_body = nullptr;
unsigned g_nthreads = 0, g_live = 0, g_bar_count = 0, g_bar_gen = 0;
bool g_progress = false;
uint8_t *g_dyn = nullptr;
size_t g_dyn_cap = 0;
long long g_clock = 0;
const char *g_kernel = "?";

void yield_to_scheduler()
{
    Fiber &f = g_fibers[g_self->linear];
    gzpb_emu_switch(&f.sp, &g_sched_sp);
}

void release_barrier_if_complete()
{
    if (g_live > 0 && g_bar_count == g_live) { g_bar_count = 0; g_bar_gen++; g_progress = true; }
}

void complete_if_ready(WarpState &W)
{
    const unsigned need = W.mask & W.exist & ~W.exited;
    if (!W.releasing && W.arrived && (W.arrived & need) == need) { W.releasing = true; g_progress = true; }
}

void fiber_main()
{
    Fiber &f = g_fibers[g_self->linear];
    (*g_body)();
    f.done = true;
    g_live--;
    g_progress = true;
    WarpState &W = g_warps[f.self.warp];
    W.exited |= 1u << f.self.lane;
    complete_if_ready(W);
    release_barrier_if_complete();
    gzpb_emu_switch(&f.sp, &g_sched_sp);
    trap("resumed a finished fiber");
}

void prepare(Fiber &f)
{
    if (!f.stack) {
        void *m = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) trap("mmap of a fiber stack failed");
        f.stack = (uint8_t *)m;
    }
    uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
    void **slot = (void **)(top - 16);
    slot[0] = (void *)&fiber_main;      // `ret` target; rsp % 16 == 8 on entry
    slot[1] = nullptr;
    void **regs = slot - 6;             // r15 r14 r13 r12 rbx rbp
    for (int i = 0; i < 6; i++) regs[i] = nullptr;
    f.sp = (void *)regs;
    f.done = false;
}
}  // namespace

[[noreturn]] void trap(const char *why)
{
    fprintf(stderr, "[gzpb_emu] kernel trap: %s", why);
    if (g_self) fprintf(stderr, " (block %u thread %u)", g_self->bid.x, g_self->tid.x);
    fprintf(stderr, "\n");
    abort();
}

void note_progress() { g_progress = true; }
void wait_yield() { yield_to_scheduler(); }
uint8_t *dyn_smem() { return g_dyn; }
long long clock() { return ++g_clock; }

void syncthreads()
{
    const unsigned gen = g_bar_gen;
    g_bar_count++;
    g_progress = true;
    release_barrier_if_complete();
    while (g_bar_gen == gen) yield_to_scheduler();
}

uint64_t collective(int op, unsigned mask, uint64_t val, uint64_t *all32)
{
    Self &me = *g_self;
    WarpState &W = g_warps[me.warp];
    const unsigned bit = 1u << me.lane;
    if (!(mask & bit)) trap("collective: the calling lane is not in its own mask");
    while (W.releasing) yield_to_scheduler();            // the previous collective is still being read
    if (W.arrived == 0) { W.mask = mask; W.op = op; }
    else if (W.mask != mask || W.op != op) trap("collective: lanes of one warp are in different collectives (divergent masks are not emulated)");
    W.val[me.lane] = val;
    W.arrived |= bit;
    g_progress = true;
    complete_if_ready(W);
    while (!W.releasing) yield_to_scheduler();
    uint64_t r = 0;
    const unsigned in = W.arrived;
    switch (op) {
    case OP_SYNCWARP: break;
    case OP_BALLOT: for (int l = 0; l < 32; l++) if (((in >> l) & 1) && W.val[l]) r |= 1ull << l; break;
    case OP_SHFL: break;
    case OP_MATCH_ANY: for (int l = 0; l < 32; l++) if (((in >> l) & 1) && W.val[l] == val) r |= 1ull << l; break;
    case OP_RED_ADD: for (int l = 0; l < 32; l++) if ((in >> l) & 1) r += W.val[l]; r &= 0xFFFFFFFFull; break;
    case OP_RED_OR: for (int l = 0; l < 32; l++) if ((in >> l) & 1) r |= W.val[l]; break;
    case OP_RED_XOR: for (int l = 0; l < 32; l++) if ((in >> l) & 1) r ^= W.val[l]; break;
    case OP_RED_AND: r = ~0ull; for (int l = 0; l < 32; l++) if ((in >> l) & 1) r &= W.val[l]; break;
    case OP_RED_MIN: r = ~0ull; for (int l = 0; l < 32; l++) if (((in >> l) & 1) && W.val[l] < r) r = W.val[l]; break;
    case OP_RED_MAX: r = 0; for (int l = 0; l < 32; l++) if (((in >> l) & 1) && W.val[l] > r) r = W.val[l]; break;
    default: trap("collective: unknown op");
    }
    if (all32) for (int l = 0; l < 32; l++) all32[l] = ((in >> l) & 1) ? W.val[l] : val;   // a lane outside the mask: undefined on a GPU
    W.departed |= bit;
    if (W.departed == W.arrived) { W.arrived = 0; W.departed = 0; W.releasing = false; g_progress = true; }
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    if (g_self) trap("nested launch");
    const unsigned n = block.x * block.y * block.z;
    if (n == 0 || n > 1024) trap("launch: bad block size");
    if (g_fibers.size() < n) g_fibers.resize(n);
    if (smem > g_dyn_cap) {
        free(g_dyn);
        g_dyn_cap = (smem + 1023) & ~(size_t)1023;
        g_dyn = (uint8_t *)aligned_alloc(1024, g_dyn_cap);
    }
    g_body = &body;
    g_nthreads = n;
    const unsigned nwarps = (n + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                if (g_dyn) memset(g_dyn, 0xA5, g_dyn_cap);      // shared memory is not zero-initialised on a GPU
                g_warps.assign(nwarps, WarpState());
                g_live = n; g_bar_count = 0; g_bar_gen = 0;
                for (unsigned t = 0; t < n; t++) {
                    Fiber &f = g_fibers[t];
                    f.self.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    f.self.bid = uint3{bx, by, bz};
                    f.self.bdim = uint3{block.x, block.y, block.z};
                    f.self.gdim = uint3{grid.x, grid.y, grid.z};
                    f.self.linear = t; f.self.lane = t & 31; f.self.warp = t >> 5;
                    g_warps[t >> 5].exist |= 1u << (t & 31);
                    prepare(f);
                }
                while (g_live > 0) {
                    g_progress = false;
                    for (unsigned t = 0; t < n; t++) {
                        Fiber &f = g_fibers[t];
                        if (f.done) continue;
                        g_self = &f.self;
                        gzpb_emu_switch(&g_sched_sp, &f.sp);
                    }
                    if (!g_progress && g_live > 0) {
                        g_self = nullptr;
                        fprintf(stderr, "[gzpb_emu] deadlock in block (%u,%u,%u): %u fibers alive, none can proceed "
                                        "(mismatched __syncthreads / warp collective / mbarrier that never completes)\n", bx, by, bz, g_live);
                        abort();
                    }
                }
                g_self = nullptr;
            }
    g_body = nullptr;
}

}  // namespace gzpb_emu

// ---- CUDA runtime subset: host memory stands in for device memory ------------------
namespace {
std::map<uintptr_t, size_t> g_pinned;
struct Fill { static void garbage(void *p, size_t n) { memset(p, 0xA5, n); } };
}

struct emu_stream { int id; };
struct emu_event { int id; };

cudaError_t cudaMalloc(void **p, size_t n)
{
    void *m = aligned_alloc(256, (n + 255) & ~(size_t)255);
    if (!m) return cudaErrorMemoryAllocation;
    Fill::garbage(m, n);          // device memory is not zero-initialised
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned)
{
    void *m = aligned_alloc(256, (n + 255) & ~(size_t)255);
    if (!m) return cudaErrorMemoryAllocation;
    Fill::garbage(m, n);
    g_pinned[(uintptr_t)m] = n;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { if (p) { g_pinned.erase((uintptr_t)p); free(p); } return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < h; r++) memmove((uint8_t *)d + r * dp, (const uint8_t *)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "gzpb_emu (CPU SIMT emulator, test only)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
    memset(a, 0, sizeof *a);
    a->type = cudaMemoryTypeUnregistered;
    auto it = g_pinned.upper_bound((uintptr_t)p);
    if (it != g_pinned.begin()) {
        --it;
        if ((uintptr_t)p < it->first + it->second) { a->type = cudaMemoryTypeHost; a->hostPointer = (void *)p; a->devicePointer = (void *)p; }
    }
    return cudaSuccess;
}
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
