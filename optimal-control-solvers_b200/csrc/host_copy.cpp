// host_copy.cpp -- the host-side memcpy of the pinned staging ring (host_stager.h).
//
// A staged device->host transfer ends with slot -> caller-array copies on the copy threads.  glibc's memcpy uses
// ordinary stores for a 4 MiB slot, so every destination line is first READ into the cache (read-for-ownership):
// three units of host-memory traffic per byte delivered, on top of the DMA that filled the slot.  With several GPUs
// downloading at once that is what saturates (round 2: pageable end-to-end 1.39x the pinned time on one GPU, 2.3x
// on two).  Non-temporal stores write the destination without reading it.  AVX2 when the CPU has it (checked at run
// time), memcpy otherwise.  Plain C++: compiled by the host compiler, no CUDA in here.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>

__attribute__((target("avx2"))) static void stream_copy_avx2(char *d, const char *s, size_t n)
{
    size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    const size_t blocks = n / 128;
    for (size_t k = 0; k < blocks; ++k) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 96), e);
        s += 128;
        d += 128;
    }
    _mm_sfence();
    memcpy(d, s, n - blocks * 128);
}

extern "C" void ocs_host_copy(void *dst, const void *src, size_t n)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= (256u << 10))
        stream_copy_avx2(static_cast<char *>(dst), static_cast<const char *>(src), n);
    else
        memcpy(dst, src, n);
}
#else
extern "C" void ocs_host_copy(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
#endif
