/*
 * orie_io.h — C ABI of the native (host, multi-threaded) reader for the
 * reference's on-disk formats: YOLOv5 label files and detector outputs.
 *
 * Replaces the per-file Python loop of lib/data.py:11-43 (load_data): for every
 * image name read `<name>.txt` if it exists, else `<name>.npy`, else no rows
 * (lib/data.py:23-28,39-41).  Text rows are split on single spaces after
 * stripping the line and parsed as float64 with correctly rounded strtod, which
 * yields the same bits as the reference's astype(float) (lib/data.py:31);
 * column 0 is the class, columns 1..4 `xc yc w h` and, for detections, the LAST
 * column of the row is the confidence (lib/data.py:33-38).  Rows of one file are
 * truncated to the shortest row's width like upstream's zip(*rows)
 * (lib/data.py:31).  Nothing here touches the GPU; liborie_io.so has no CUDA
 * dependency.
 *
 * Files this reader does not want to judge (a token that is not a plain decimal
 * / exponent / inf / nan literal, an empty token from a doubled space, an .npy
 * that is not a C-ordered little-endian f4/f8 matrix, fewer columns than needed)
 * are NOT errors here: the image is reported in the fallback list with zero rows
 * and the caller re-reads exactly that file with the reference-equivalent Python
 * path, so error behaviour (ValueError from astype(float)) stays upstream's.
 */
#ifndef ORIE_IO_H
#define ORIE_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orie_rows orie_rows_t;

enum { ORIE_IO_OK = 0, ORIE_IO_EINVAL = 1, ORIE_IO_ENOMEM = 2, ORIE_IO_EIO = 3 };

const char *orie_io_last_error(void);

/*
 * Read the rows of `count` images from directory `dir`; names[i] is the image
 * name without extension (lib/data.py:54-56).  with_conf != 0: detections, 6
 * output columns `cls xc yc w h conf`; otherwise labels, 5 columns.
 * threads <= 0: one per online CPU (capped at 64).
 */
int orie_io_read_rows(const char *dir, const char *const *names, int64_t count, int with_conf, int threads,
                      orie_rows_t **out);

int64_t orie_io_num_images(const orie_rows_t *r);
int64_t orie_io_num_rows(const orie_rows_t *r);
int orie_io_num_cols(const orie_rows_t *r);
const int64_t *orie_io_offsets(const orie_rows_t *r);     /* int64[count + 1], CSR by image */
const double *orie_io_data(const orie_rows_t *r);         /* f64[num_rows][num_cols], row-major */
/* images whose file must be re-read by the caller (ascending); they have zero rows here */
int64_t orie_io_num_fallback(const orie_rows_t *r);
const int64_t *orie_io_fallback(const orie_rows_t *r);
void orie_io_free(orie_rows_t *r);

#ifdef __cplusplus
}
#endif
#endif /* ORIE_IO_H */
