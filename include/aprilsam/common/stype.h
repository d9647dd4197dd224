/* stype.h -- self-describing serialisation of the AprilSAM graph files (".graph").
 *
 * Replaces aprilsam/common/stype.h + encode_bytes.h of the reference (same public names and
 * struct layout, so that callers can register their own attribute types).  On-disk format
 * (reference: common/stype.c:75-107, common/encode_bytes.h): every object is framed as
 *     u64 magic | u32 len, name bytes | u32 payload length | payload | u64 magic
 * all integers and IEEE doubles big-endian; a NULL object has an empty name and length 0.
 */
#ifndef ASAM_STYPE_H
#define ASAM_STYPE_H

#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stype stype_t;
struct stype { /* reference: common/stype.h:36-57 */
    char *name;
    /* data == NULL: dry run, only *datapos advances */
    void (*encode)(const stype_t *stype, uint8_t *data, uint32_t *datapos, const void *obj);
    void *(*decode)(const stype_t *stype, const uint8_t *data, uint32_t *datapos, uint32_t datalen);
    void *(*copy)(const stype_t *stype, const void *obj);
    void (*destroy)(const stype_t *stype, void *obj);
    void *impl;
};

void stype_register_basic_types(void); /* "uint64", "string" (common/stype_basic_types.c) */
void stype_register(const stype_t *stype);
stype_t *stype_get(char *name);
void stype_encode_object(uint8_t *data, uint32_t *datapos, const stype_t *stype, const void *obj);
void *stype_decode_object(const uint8_t *data, uint32_t *datapos, uint32_t datalen, const stype_t **outstype);
int stype_write_file(const stype_t *stype, void *obj, const char *path); /* 0 on success */
void *stype_read_file(const char *path);

/* big-endian primitives (common/encode_bytes.h); out == NULL only advances *outpos */
void encode_u8(uint8_t *out, uint32_t *outpos, uint8_t v);
void encode_u32(uint8_t *out, uint32_t *outpos, uint32_t v);
void encode_u64(uint8_t *out, uint32_t *outpos, uint64_t v);
void encode_f64(uint8_t *out, uint32_t *outpos, double v);
void encode_string_u32(uint8_t *out, uint32_t *outpos, const char *s);
uint8_t decode_u8(const uint8_t *in, uint32_t *inpos, uint32_t inlen);
uint32_t decode_u32(const uint8_t *in, uint32_t *inpos, uint32_t inlen);
uint64_t decode_u64(const uint8_t *in, uint32_t *inpos, uint32_t inlen);
double decode_f64(const uint8_t *in, uint32_t *inpos, uint32_t inlen);
char *decode_string_u32(const uint8_t *in, uint32_t *inpos, uint32_t inlen); /* malloc'd */

#ifdef __cplusplus
}
#endif
#endif
