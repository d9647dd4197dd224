/* matd.h -- tiny dense row-major double matrix (aprilsam_b200).
 *
 * Boundary type: factors carry their information matrix as `matd_t *W`
 * (reference: aprilsam/common/matd.h:46-51, {unsigned nrows, ncols; double data[];},
 * element (r,c) at data[r*ncols+c]).  Only what callers of the solver path need is
 * provided; the reference's expression evaluator / SVD / LU are out of scope -- the
 * 3x3 products of the solver run inside the CUDA kernels.
 */
#ifndef ASAM_MATD_H
#define ASAM_MATD_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    unsigned int nrows, ncols;
    double data[];
} matd_t;

#define MATD_EL(m, row, col) (m)->data[((row) * (m)->ncols + (col))]

matd_t *matd_create(int rows, int cols);                          /* zero-filled */
matd_t *matd_create_data(int rows, int cols, const double *data); /* row-major copy */
matd_t *matd_identity(int dim);
matd_t *matd_copy(const matd_t *m);
void matd_destroy(matd_t *m);
void matd_print(const matd_t *m, const char *fmt); /* common/matd.h:158 */

#ifdef __cplusplus
}
#endif
#endif
