/* cliopt.c -- option parser, clock and matrix print used by the example programs (host C, no GPU).
 * with these the reference's four example programs compile and link against this
 * library unchanged (tests/test_host_cpu.py::test_reference_examples_link_unchanged). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "asam_host.h"
#include "common/getopt.h"

enum { OPT_BOOL, OPT_INT, OPT_STRING, OPT_DOUBLE, OPT_SPACER };

typedef struct {
    int kind;
    char sopt;
    char *lname, *help, *value; /* value: current value as text */
    int specified;
} opt_t;

struct getopt {
    opt_t *opt;
    int n, cap;
    zarray_t *extra;
};

ASAM_API getopt_t *getopt_create(void)
{
    getopt_t *g = calloc(1, sizeof(*g));
    g->extra = zarray_create(sizeof(char *));
    return g;
}

ASAM_API void getopt_destroy(getopt_t *g)
{
    if (!g)
        return;
    for (int i = 0; i < g->n; i++) {
        free(g->opt[i].lname);
        free(g->opt[i].help);
        free(g->opt[i].value);
    }
    free(g->opt);
    for (int i = 0; i < zarray_size(g->extra); i++) {
        char *s;
        zarray_get(g->extra, i, &s);
        free(s);
    }
    zarray_destroy(g->extra);
    free(g);
}

static void add(getopt_t *g, int kind, char sopt, const char *lname, const char *def, const char *help)
{
    if (g->n == g->cap) {
        g->cap = g->cap ? 2 * g->cap : 8;
        g->opt = realloc(g->opt, sizeof(opt_t) * (size_t) g->cap);
    }
    opt_t *o = &g->opt[g->n++];
    o->kind = kind;
    o->sopt = sopt;
    o->lname = strdup(lname ? lname : "");
    o->help = strdup(help ? help : "");
    o->value = strdup(def ? def : "");
    o->specified = 0;
}

ASAM_API void getopt_add_spacer(getopt_t *g, const char *s) { add(g, OPT_SPACER, 0, "", "", s); }
ASAM_API void getopt_add_bool(getopt_t *g, char sopt, const char *lname, int def, const char *help)
{
    add(g, OPT_BOOL, sopt, lname, def ? "true" : "false", help);
}
ASAM_API void getopt_add_int(getopt_t *g, char sopt, const char *lname, const char *def, const char *help)
{
    add(g, OPT_INT, sopt, lname, def, help);
}
ASAM_API void getopt_add_string(getopt_t *g, char sopt, const char *lname, const char *def, const char *help)
{
    add(g, OPT_STRING, sopt, lname, def, help);
}
ASAM_API void getopt_add_double(getopt_t *g, char sopt, const char *lname, const char *def, const char *help)
{
    add(g, OPT_DOUBLE, sopt, lname, def, help);
}

static opt_t *find_long(getopt_t *g, const char *name, size_t len)
{
    for (int i = 0; i < g->n; i++)
        if (g->opt[i].kind != OPT_SPACER && strlen(g->opt[i].lname) == len && !strncmp(g->opt[i].lname, name, len))
            return &g->opt[i];
    return NULL;
}

static opt_t *find_short(getopt_t *g, char c)
{
    for (int i = 0; i < g->n; i++)
        if (g->opt[i].kind != OPT_SPACER && g->opt[i].sopt && g->opt[i].sopt == c)
            return &g->opt[i];
    return NULL;
}

static void set_value(opt_t *o, const char *v)
{
    free(o->value);
    o->value = strdup(v);
    o->specified = 1;
}

ASAM_API int getopt_parse(getopt_t *g, int argc, char *argv[], int showErrors)
{
    int ok = 1;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        opt_t *o = NULL;
        const char *inline_val = NULL;
        if (!strncmp(a, "--", 2) && a[2]) {
            const char *eq = strchr(a + 2, '=');
            o = find_long(g, a + 2, eq ? (size_t) (eq - (a + 2)) : strlen(a + 2));
            inline_val = eq ? eq + 1 : NULL;
        } else if (a[0] == '-' && a[1] && !a[2]) {
            o = find_short(g, a[1]);
        } else {
            char *s = strdup(a);
            zarray_add(g->extra, &s);
            continue;
        }
        if (!o) {
            if (showErrors)
                printf("Unknown option %s\n", a);
            ok = 0;
            continue;
        }
        if (o->kind == OPT_BOOL) {
            /* "--flag", "--flag=false", or "--flag true|false" */
            if (inline_val)
                set_value(o, inline_val);
            else if (i + 1 < argc && (!strcmp(argv[i + 1], "true") || !strcmp(argv[i + 1], "false")))
                set_value(o, argv[++i]);
            else
                set_value(o, "true");
        } else if (inline_val) {
            set_value(o, inline_val);
        } else if (i + 1 < argc) {
            set_value(o, argv[++i]);
        } else {
            if (showErrors)
                printf("Option %s requires a value\n", a);
            ok = 0;
        }
    }
    return ok;
}

ASAM_API char *getopt_get_usage(getopt_t *g)
{
    size_t cap = 256;
    for (int i = 0; i < g->n; i++)
        cap += strlen(g->opt[i].lname) + strlen(g->opt[i].help) + strlen(g->opt[i].value) + 64;
    char *out = malloc(cap);
    size_t pos = 0;
    pos += (size_t) snprintf(out + pos, cap - pos, "Usage: [options]\n\n");
    for (int i = 0; i < g->n; i++) {
        const opt_t *o = &g->opt[i];
        if (o->kind == OPT_SPACER) {
            pos += (size_t) snprintf(out + pos, cap - pos, "\n%s\n", o->help);
            continue;
        }
        char sh[8] = "   ";
        if (o->sopt)
            snprintf(sh, sizeof(sh), "-%c ", o->sopt);
        pos += (size_t) snprintf(out + pos, cap - pos, " %s| --%-20s [ %-10s ]   %s\n", sh, o->lname, o->value, o->help);
    }
    return out;
}

ASAM_API void getopt_do_usage(getopt_t *g)
{
    char *u = getopt_get_usage(g);
    fputs(u, stdout);
    free(u);
}

static const opt_t *need(getopt_t *g, const char *lname, int kind)
{
    const opt_t *o = find_long(g, lname, strlen(lname));
    if (!o || o->kind != kind)
        asam_fatal("getopt: no %s option named '%s'", kind == OPT_BOOL ? "bool" : kind == OPT_INT ? "int" : kind == OPT_DOUBLE ? "double" : "string", lname);
    return o;
}

ASAM_API const char *getopt_get_string(getopt_t *g, const char *lname) { return need(g, lname, OPT_STRING)->value; }
ASAM_API int getopt_get_int(getopt_t *g, const char *lname) { return (int) strtol(need(g, lname, OPT_INT)->value, NULL, 10); }
ASAM_API double getopt_get_double(getopt_t *g, const char *lname) { return strtod(need(g, lname, OPT_DOUBLE)->value, NULL); }
ASAM_API int getopt_get_bool(getopt_t *g, const char *lname)
{
    const char *v = need(g, lname, OPT_BOOL)->value;
    return !strcmp(v, "true") || !strcmp(v, "1");
}
ASAM_API int getopt_was_specified(getopt_t *g, const char *lname)
{
    const opt_t *o = find_long(g, lname, strlen(lname));
    return o ? o->specified : 0;
}
ASAM_API const zarray_t *getopt_get_extra_args(getopt_t *g) { return g->extra; }

/* microseconds since the epoch (common/time_util.c) */
ASAM_API int64_t utime_now(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (int64_t) tv.tv_sec * 1000000 + tv.tv_usec;
}

/* every entry with `fmt`, one row per line (common/matd.c) */
ASAM_API void matd_print(const matd_t *m, const char *fmt)
{
    for (unsigned i = 0; i < m->nrows; i++) {
        for (unsigned j = 0; j < m->ncols; j++)
            printf(fmt, MATD_EL(m, i, j));
        printf("\n");
    }
}
