/* zarray.h -- growable array of fixed-size elements (aprilsam_b200).
 *
 * Boundary type: `graph->nodes` / `graph->factors` are zarray_t of POINTERS that callers
 * fill with zarray_add() (reference: aprilsam/common/zarray.h:44-51, struct layout
 * {size_t el_sz; int size; int alloc; char *data;} = 24 bytes on x86-64).  Only the
 * layout and the call names are shared with the reference; the code below is ours.
 */
#ifndef ASAM_ZARRAY_H
#define ASAM_ZARRAY_H

#include <assert.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zarray zarray_t;
struct zarray {
    size_t el_sz; /* bytes per element           */
    int size;     /* elements in use             */
    int alloc;    /* elements allocated          */
    char *data;
};

static inline zarray_t *zarray_create(size_t el_sz)
{
    zarray_t *za = (zarray_t *) calloc(1, sizeof(zarray_t));
    za->el_sz = el_sz;
    return za;
}

static inline void zarray_destroy(zarray_t *za)
{
    if (!za)
        return;
    free(za->data);
    free(za);
}

static inline int zarray_size(const zarray_t *za) { return za->size; }

static inline void zarray_ensure_capacity(zarray_t *za, int capacity)
{
    if (capacity <= za->alloc)
        return;
    int want = za->alloc > 0 ? za->alloc : 8;
    while (want < capacity)
        want *= 2;
    za->data = (char *) realloc(za->data, (size_t) want * za->el_sz);
    za->alloc = want;
}

static inline void zarray_add(zarray_t *za, const void *p)
{
    zarray_ensure_capacity(za, za->size + 1);
    memcpy(za->data + (size_t) za->size * za->el_sz, p, za->el_sz);
    za->size++;
}

static inline void zarray_get(const zarray_t *za, int idx, void *p)
{
    assert(idx >= 0 && idx < za->size);
    memcpy(p, za->data + (size_t) idx * za->el_sz, za->el_sz);
}

static inline void zarray_get_volatile(const zarray_t *za, int idx, void *p)
{
    assert(idx >= 0 && idx < za->size);
    *((void **) p) = za->data + (size_t) idx * za->el_sz;
}

static inline void zarray_set(zarray_t *za, int idx, const void *p, void *outp)
{
    assert(idx >= 0 && idx < za->size);
    char *slot = za->data + (size_t) idx * za->el_sz;
    if (outp)
        memcpy(outp, slot, za->el_sz);
    memcpy(slot, p, za->el_sz);
}

static inline void zarray_clear(zarray_t *za) { za->size = 0; }

#ifdef __cplusplus
}
#endif
#endif
