// mono_alloc.cc - a monotonic operator new/delete for ONE build variant of the reference extractor
// (oracle/_ref/libref_orb_mono.so; test infrastructure only).
//
// Why: DistributeOctTree sorts pair<int, ExtractorNode*> (reference src/ORBextractor.cc:684), so nodes holding the
// same number of keys are split in order of their HEAP ADDRESS.  Under glibc malloc that order depends on the free-list
// history; under an allocator that never reuses memory, addresses grow with creation order, which is exactly the
// canonical rule the oracle and the CUDA path adopt (DESIGN.md rule 1: the later-created node is split first).  Linking
// this file (with -Wl,-Bsymbolic, so that only the library's own allocations are affected) turns the reference binary
// into a deterministic one that must then agree with the oracle byte for byte.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

namespace {
const size_t ARENA_BYTES = (size_t)16 << 30; // virtual only: MAP_NORESERVE, pages are committed when touched
const size_t BIG = (size_t)64 << 10;         // image buffers go to malloc: only small (list-node) order matters
char *g_base = nullptr;
std::atomic<size_t> g_off{0};
std::atomic<int> g_init{0};

char *arena()
{
    if (g_init.load(std::memory_order_acquire) == 2) return g_base;
    int expect = 0;
    if (g_init.compare_exchange_strong(expect, 1)) {
        void *p = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) std::abort();
        g_base = static_cast<char *>(p);
        g_init.store(2, std::memory_order_release);
    } else {
        while (g_init.load(std::memory_order_acquire) != 2) {}
    }
    return g_base;
}
void *bump(size_t n)
{
    if (n >= BIG) { void *p = std::malloc(n); if (!p) throw std::bad_alloc(); return p; }
    char *b = arena();
    const size_t sz = (n + 15) & ~(size_t)15;
    const size_t at = g_off.fetch_add(sz ? sz : 16);
    if (at + sz > ARENA_BYTES) std::abort();
    return b + at;
}
void release(void *p)
{
    if (!p) return;
    const char *c = static_cast<const char *>(p);
    if (g_base && c >= g_base && c < g_base + ARENA_BYTES) return; // never reused
    std::free(p);
}
} // namespace

void *operator new(size_t n) { return bump(n); }
void *operator new[](size_t n) { return bump(n); }
void operator delete(void *p) noexcept { release(p); }
void operator delete[](void *p) noexcept { release(p); }
void operator delete(void *p, size_t) noexcept { release(p); }
void operator delete[](void *p, size_t) noexcept { release(p); }

extern "C" size_t ref_mono_bytes() { return g_off.load(); }
