/* par_for.h -- tiny pthread work-sharing loop for the CPU oracle
 * (test infrastructure only; the image has no OpenMP runtime). */
#ifndef RO_PAR_FOR_H
#define RO_PAR_FOR_H
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

typedef void (*ro_body_fn)(int i, void *ctx);
typedef struct { int n; volatile int *next; ro_body_fn f; void *ctx; } ro_pf_t;

static void *ro_pf_worker(void *v)
{
    ro_pf_t *p = (ro_pf_t *)v;
    for (;;) {
        int i = __sync_fetch_and_add((int *)p->next, 1);
        if (i >= p->n) break;
        p->f(i, p->ctx);
    }
    return NULL;
}

static void ro_par_for(int n, int threads, ro_body_fn f, void *ctx)
{
    if (threads <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); threads = c > 0 ? (int)c : 1; }
    if (threads > n) threads = n;
    if (threads <= 1) { for (int i = 0; i < n; i++) f(i, ctx); return; }
    volatile int next = 0;
    ro_pf_t p = {n, &next, f, ctx};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, ro_pf_worker, &p);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}
#endif
