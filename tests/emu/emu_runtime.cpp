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
    long bar_wait = -1;      // generation of the CTA barrier this fiber sleeps on (-1 = runnable): the scheduler skips sleepers
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
struct TmaPending { void *bar; void *dst; const void *src; uint32_t bytes; };
std::vector<TmaPending> g_tma_pending;       // bulk copies issued, not yet waited for (tma_late)
unsigned long long g_launches_queued = 0;    // kernels launched in this process (gzpb_emu_kernel_launches)
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
    Fiber &f = g_fibers[g_self->linear];
    while (g_bar_gen == gen) { f.bar_wait = (long)gen; yield_to_scheduler(); }
    f.bar_wait = -1;
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
    if (smem != g_dyn_cap) {                       // exact size: an access past the launch's dynamic shared memory is an ASan error
        free(g_dyn);
        g_dyn = nullptr;
        g_dyn_cap = smem;
        if (smem) { void *m = nullptr; if (posix_memalign(&m, 1024, smem) != 0) trap("out of memory for dynamic shared memory"); g_dyn = (uint8_t *)m; }
    }
    g_body = &body;
    g_nthreads = n;
    static const int sched = [] { const char *e = getenv("GZPB_EMU_SCHED"); return !e ? 0 : !strcmp(e, "reverse") ? 1 : !strncmp(e, "random", 6) ? 2 : 0; }();
    static uint64_t rng_state = [] { const char *e = getenv("GZPB_EMU_SCHED"); const char *c = e ? strchr(e, ':') : nullptr; return c ? strtoull(c + 1, nullptr, 10) * 2 + 1 : 0x9E3779B97F4A7C15ull; }();
    auto sched_rng = [&]() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; };
    unsigned rot = 0;
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
                    f.self.linear = t; f.self.lane = t & 31; f.self.warp = t >> 5; f.bar_wait = -1;
                    g_warps[t >> 5].exist |= 1u << (t & 31);
                    prepare(f);
                }
                while (g_live > 0) {
                    g_progress = false;
                    for (unsigned i = 0; i < n; i++) {
                        // scheduling order between rendezvous points: ascending thread index by default; reversed or
                        // pseudo-random per pass on request — a result that depends on it is a missing barrier
                        unsigned t = i;
                        if (sched == 1) t = n - 1 - i;
                        else if (sched == 2) { if (i == 0) rot = (unsigned)(sched_rng() % n); t = (i + rot) % n; if (rot & 1) t = n - 1 - t; }
                        Fiber &f = g_fibers[t];
                        if (f.done) continue;
                        if (f.bar_wait >= 0 && (unsigned)f.bar_wait == g_bar_gen) continue;   // still asleep on __syncthreads
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
                if (!g_tma_pending.empty()) {
                    fprintf(stderr, "[gzpb_emu] block (%u,%u,%u) ended with %zu bulk copies in flight\n", bx, by, bz, g_tma_pending.size());
                    abort();
                }
            }
    g_body = nullptr;
}

bool tma_late() { static const bool late = [] { const char *e = getenv("GZPB_EMU_TMA"); return !(e && !strcmp(e, "eager")); }(); return late; }
void tma_defer(void *bar, void *dst, const void *src, uint32_t bytes) { g_tma_pending.push_back(TmaPending{bar, dst, src, bytes}); }
uint32_t tma_deliver(void *bar)
{
    uint32_t moved = 0;
    for (size_t i = 0; i < g_tma_pending.size();) {
        if (g_tma_pending[i].bar == bar) {
            memcpy(g_tma_pending[i].dst, g_tma_pending[i].src, g_tma_pending[i].bytes);
            moved += g_tma_pending[i].bytes;
            g_tma_pending.erase(g_tma_pending.begin() + (long)i);
        } else i++;
    }
    return moved;
}

}  // namespace gzpb_emu

// ---- CUDA runtime subset: host memory stands in for device memory ------------------
//
// Streams are LAZY by default: every asynchronous call (cudaMemcpyAsync between device and pinned
// memory, cudaMemsetAsync, kernel launch, event record / wait) is queued on its stream and executed
// only when the host synchronises with it (cudaEventSynchronize, cudaStreamSynchronize,
// cudaDeviceSynchronize, cudaFree*, a blocking copy) — and then only as far as that synchronisation
// requires.  Host code that reuses a pinned source before its copy ran, or reads a result before
// waiting for it, therefore produces wrong bytes here just as it would (sometimes) on a GPU.
// GZPB_EMU_EAGER=1 restores immediate execution; GZPB_EMU_DEVICES=n exposes n devices.
#include <deque>
#include <set>

namespace {
std::map<uintptr_t, size_t> g_pinned, g_device;
struct Fill { static void garbage(void *p, size_t n) { memset(p, 0xA5, n); } };
int g_cur_device = 0;
bool env_flag(const char *name) { const char *v = getenv(name); return v && *v && *v != '0'; }
int env_devices() { const char *v = getenv("GZPB_EMU_DEVICES"); int n = v ? atoi(v) : 1; return n < 1 ? 1 : n > 16 ? 16 : n; }
bool lazy() { static const bool l = !env_flag("GZPB_EMU_EAGER"); return l; }
bool in_map(const std::map<uintptr_t, size_t> &m, const void *p)
{
    auto it = m.upper_bound((uintptr_t)p);
    if (it == m.begin()) return false;
    --it;
    return (uintptr_t)p < it->first + it->second;
}
bool pageable(const void *p) { return !in_map(g_pinned, p) && !in_map(g_device, p); }
}  // namespace

struct emu_event { uint64_t recorded = 0, completed = 0; emu_stream *last = nullptr; int device = 0; };
struct emu_op { int kind; std::function<void()> fn; emu_event *ev; uint64_t target; };   // 0 work, 1 record, 2 wait
struct emu_stream { std::deque<emu_op> q; int device = 0; };

namespace {
std::set<emu_stream *> g_streams;
unsigned long long g_deferred = 0;
emu_stream *legacy_stream()
{
    static emu_stream *s = nullptr;
    if (!s) { s = new emu_stream(); g_streams.insert(s); }
    return s;
}
emu_stream *S(cudaStream_t s) { return s ? s : legacy_stream(); }

void complete_event(emu_event *e, uint64_t target, int depth);

// executes the head op of `s`; a wait first forces the awaited record (on whichever stream holds it)
void step(emu_stream *s, int depth)
{
    if (depth > 64) gzpb_emu::trap("emulated streams: event wait cycle");
    emu_op op = std::move(s->q.front());
    s->q.pop_front();
    if (op.kind == 0) op.fn();
    else if (op.kind == 1) { if (op.ev->completed < op.target) op.ev->completed = op.target; }
    else complete_event(op.ev, op.target, depth + 1);
}

void complete_event(emu_event *e, uint64_t target, int depth)
{
    while (e->completed < target) {
        emu_stream *s = e->last;
        if (!s || s->q.empty()) gzpb_emu::trap("emulated streams: waiting for an event whose record is not queued anywhere");
        step(s, depth);
    }
}

void drain(emu_stream *s) { while (!s->q.empty()) step(s, 0); }
void drain_all() { for (emu_stream *s : g_streams) drain(s); }

void enqueue(cudaStream_t st, std::function<void()> fn)
{
    if (!lazy()) { fn(); return; }
    g_deferred++;
    S(st)->q.push_back(emu_op{0, std::move(fn), nullptr, 0});
}
}  // namespace

// test hook: how many operations were queued on lazy streams so far (0 in eager mode)
extern "C" unsigned long long gzpb_emu_deferred_ops() { return g_deferred; }

namespace gzpb_emu {
void launch_async(void *stream, dim3 grid, dim3 block, size_t smem, std::function<void()> body)
{
    emu_stream *s = S((cudaStream_t)stream);
    if (s->device != g_cur_device && s != legacy_stream()) trap("kernel launch on a stream of another device (missing cudaSetDevice)");
    g_launches_queued++;
    enqueue((cudaStream_t)stream, [=]() { launch(grid, block, smem, body); });
}
}  // namespace gzpb_emu

// kernels launched so far in this process (tests compare it with the library's own gzpb_launch_count)
extern "C" unsigned long long gzpb_emu_kernel_launches() { return gzpb_emu::g_launches_queued; }

cudaError_t cudaMalloc(void **p, size_t n)
{
    void *m = nullptr;            // exact size: under AddressSanitizer (GZPB_EMU_ASAN=1) the redzone starts right after byte n
    if (posix_memalign(&m, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    Fill::garbage(m, n);          // device memory is not zero-initialised
    g_device[(uintptr_t)m] = n ? n : 1;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { if (p) { drain_all(); g_device.erase((uintptr_t)p); free(p); } return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned)
{
    void *m = nullptr;
    if (posix_memalign(&m, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    Fill::garbage(m, n);
    g_pinned[(uintptr_t)m] = n ? n : 1;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { if (p) { drain_all(); g_pinned.erase((uintptr_t)p); free(p); } return cudaSuccess; }
// blocking copies run on the legacy stream: everything queued before them is visible to them
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { drain_all(); memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t st)
{
    if (pageable(d) || pageable(s)) {       // pageable host memory: the call returns after the bytes were staged / delivered
        drain(S(st));
        memmove(d, s, n);
        return cudaSuccess;
    }
    enqueue(st, [=]() { memmove(d, s, n); });
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t st)
{
    auto fn = [=]() { for (size_t r = 0; r < h; r++) memmove((uint8_t *)d + r * dp, (const uint8_t *)s + r * sp, w); };
    if (pageable(d) || pageable(s)) { drain(S(st)); fn(); return cudaSuccess; }
    enqueue(st, fn);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *d, int v, size_t n) { drain_all(); memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) { enqueue(st, [=]() { memset(d, v, n); }); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= env_devices()) return cudaErrorInvalidValue; g_cur_device = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = g_cur_device; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = env_devices(); return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "gzpb_emu (CPU SIMT emulator, test only)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { drain_all(); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream(); (*s)->device = g_cur_device; g_streams.insert(*s); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { if (s) { drain(s); g_streams.erase(s); delete s; } return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { drain(S(s)); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned)
{
    if (!lazy() || e->recorded == 0) return cudaSuccess;
    S(s)->q.push_back(emu_op{2, nullptr, e, e->recorded});
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event(); (*e)->device = g_cur_device; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { if (e) { if (e->completed < e->recorded) complete_event(e, e->recorded, 0); delete e; } return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    emu_stream *st = S(s);
    if (st != legacy_stream() && st->device != e->device) return cudaErrorInvalidValue;   // event and stream must share a device
    e->recorded++;
    if (!lazy()) { e->completed = e->recorded; return cudaSuccess; }
    e->last = st;
    st->q.push_back(emu_op{1, nullptr, e, e->recorded});
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { complete_event(e, e->recorded, 0); return cudaSuccess; }
// a query lets "time pass": the stream holding the pending record advances by ONE operation, so polling
// loops terminate while a query right after a submit still sees cudaErrorNotReady
cudaError_t cudaEventQuery(cudaEvent_t e)
{
    if (e->completed >= e->recorded) return cudaSuccess;
    if (e->last && !e->last->q.empty()) step(e->last, 0);
    return e->completed >= e->recorded ? cudaSuccess : cudaErrorNotReady;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    complete_event(a, a->recorded, 0); complete_event(b, b->recorded, 0);
    *ms = 0.0f;
    return cudaSuccess;
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
    memset(a, 0, sizeof *a);
    a->type = cudaMemoryTypeUnregistered;
    if (in_map(g_pinned, p)) { a->type = cudaMemoryTypeHost; a->hostPointer = (void *)p; a->devicePointer = (void *)p; }
    else if (in_map(g_device, p)) { a->type = cudaMemoryTypeDevice; a->devicePointer = (void *)p; }
    return cudaSuccess;
}
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
