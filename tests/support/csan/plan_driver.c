/* plan_driver.c -- sanitizer driver of the host symbolic layer (tests/test_host_cpu.py::test_host_plan_under_sanitizers):
 * builds the plan of a pseudo-random pose graph (optionally the schedule of every rank of a multi-GPU job), replays 40
 * incremental appends, frees everything.  Compiled with -fsanitize=address,undefined together with plan.c + ordering.c and
 * device_stubs.c (the device entry points are never reached with dev == NULL). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "asam_host.h"
void asam_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void asam_fatal(const char *fmt, ...) { (void) fmt; abort(); }
static unsigned rs = 12345;
static int rnd(int n) { rs = rs * 1103515245u + 12345u; return (int) ((rs >> 8) % (unsigned) n); }
int main(int argc, char **argv)
{
    int N = argc > 1 ? atoi(argv[1]) : 3000, world = argc > 2 ? atoi(argv[2]) : 1;
    int cap = 6 * N, F = 0;
    int *ft = malloc(sizeof(int) * cap), *fa = malloc(sizeof(int) * cap), *fb = malloc(sizeof(int) * cap);
    ft[F] = 2; fa[F] = 0; fb[F] = -1; F++;
    int N0 = N - 40;
    for (int i = 1; i < N0; i++) {
        ft[F] = 1; fa[F] = i - 1; fb[F] = i; F++;
        if (i > 12 && rnd(3) == 0) { ft[F] = 1; fa[F] = rnd(i - 10); fb[F] = i; F++; }
        if (i > 50 && rnd(10) == 0) { ft[F] = 1; fa[F] = i; fb[F] = i - 1 - rnd(40); F++; }
    }
    for (int r = 0; r < world; r++) {
        plan_t pl;
        memset(&pl, 0, sizeof(pl));
        pl.world = world; pl.rank = r;
        if (plan_build(&pl, NULL, N0, F, ft, fa, fb)) { fprintf(stderr, "build failed\n"); return 1; }
        printf("rank %d/%d: nsn %d levels %d ntasks %d leaf %d top %d shards %d max_m %d\n", r, world, pl.nsn, pl.n_levels, pl.ntasks, pl.n_leaf, pl.n_top, pl.n_shards, pl.max_m);
        if (world == 1) { /* incremental appends */
            int F2 = F, Ncur = N0;
            for (int k = N0; k < N; k++) {
                int Fprev = F2;
                ft[F2] = 1; fa[F2] = k - 1; fb[F2] = k; F2++;
                if (rnd(2)) { ft[F2] = 1; fa[F2] = rnd(k - 10); fb[F2] = k; F2++; }
                /* marked = root paths of the old endpoints */
                int *marked = malloc(sizeof(int) * (size_t) (k + 1)), nm = 0; char *seen = calloc((size_t) k + 1, 1);
                for (int f = Fprev; f < F2; f++) {
                    int ends[2] = { fa[f], fb[f] };
                    for (int e = 0; e < 2; e++) {
                        int v = ends[e];
                        while (v >= 0 && v < Ncur && !seen[v]) {
                            seen[v] = 1; marked[nm++] = v;
                            int pp = pl.parent_pos[pl.pos[v]];
                            v = pp >= 0 ? pl.order[pp] : -1;
                        }
                    }
                }
                int *tasks = NULL, *nwait = NULL, nt = 0;
                int rc = plan_append(&pl, NULL, k + 1, F2, ft, fa, fb, marked, nm, &tasks, &nwait, NULL, &nt);
                if (rc) { fprintf(stderr, "append rc %d at %d\n", rc, k); return 1; }
                free(tasks); free(nwait); free(marked); free(seen);
                Ncur = k + 1;
            }
            printf("after appends: N %d nsn %d max_m %d\n", pl.N, pl.nsn, pl.max_m);
        }
        plan_free(&pl);
    }
    free(ft); free(fa); free(fb);
    return 0;
}
