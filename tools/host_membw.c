/* Host memory bandwidth of the bench box, the other ceiling of skyjo_step_host (the host fills 127 B per env-step):
 * T threads each write (plain stores / non-temporal stores) or copy a private 64 MB buffer.
 *   gcc -O2 -pthread -mavx2 tools/host_membw.c -o /tmp/host_membw && /tmp/host_membw <threads> */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define BYTES (64u << 20)
#define REPS 8
static int g_mode;
static pthread_barrier_t bar;
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *work(void *arg) {
    uint8_t *a = aligned_alloc(64, BYTES), *b = aligned_alloc(64, BYTES);
    memset(a, 1, BYTES); memset(b, 2, BYTES);
    pthread_barrier_wait(&bar);
    for (int r = 0; r < REPS; ++r) {
        if (g_mode == 0) memset(a, r, BYTES);
        else if (g_mode == 1) {
            __m256i v = _mm256_set1_epi8((char)r);
            for (size_t i = 0; i < BYTES; i += 32) _mm256_stream_si256((__m256i *)(a + i), v);
            _mm_sfence();
        } else memcpy(a, b, BYTES);
    }
    pthread_barrier_wait(&bar);
    *(double *)arg = a[123] + b[5];
    free(a); free(b);
    return 0;
}

int main(int argc, char **argv) {
    int T = argc > 1 ? atoi(argv[1]) : 4;
    const char *names[3] = {"plain stores (memset)", "non-temporal stores", "memcpy (read + write)"};
    for (g_mode = 0; g_mode < 3; ++g_mode) {
        pthread_t th[256]; double sink[256];
        pthread_barrier_init(&bar, 0, T + 1);
        for (int i = 0; i < T; ++i) pthread_create(&th[i], 0, work, &sink[i]);
        pthread_barrier_wait(&bar);
        double t0 = now();
        pthread_barrier_wait(&bar);
        double dt = now() - t0;
        for (int i = 0; i < T; ++i) pthread_join(th[i], 0);
        pthread_barrier_destroy(&bar);
        printf("%d threads, %s: %.1f GB/s written\n", T, names[g_mode], (double)T * BYTES * REPS / dt / 1e9);
    }
    return 0;
}
