/* linne_b200.h -- extension entry points of liblinne_b200.so next to the reference API.
 *
 * These do not exist in the reference; they expose what a GPU implementation adds: which device a
 * handle runs on, launch accounting for benchmarks, and (later rounds) device-resident batch entry
 * points.  Plain C ABI: pointers and sizes only.
 */
#ifndef LINNE_B200_H_INCLUDED
#define LINNE_B200_H_INCLUDED

#include "linne.h"
#include "linne_encoder.h"
#include "linne_decoder.h"

#ifdef __cplusplus
extern "C" {
#endif

/* "cuda-sm_100a" for the product library. */
const char *LINNEB200_Backend(void);

/* 1 if a CUDA device is usable from this process, else 0 (handles cannot be created then). */
int LINNEB200_DeviceAvailable(void);

/* Kernels launched so far on behalf of a handle (for bench.py's gpu_launches accounting). */
uint64_t LINNEB200_EncoderLaunchCount(const struct LINNEEncoder *encoder);
uint64_t LINNEB200_DecoderLaunchCount(const struct LINNEDecoder *decoder);

/* Run a handle's GPU work on an externally owned CUDA stream (cudaStream_t passed as void*), so a
 * caller can bracket the work with its own CUDA events.  API calls stay synchronous. */
void LINNEB200_EncoderUseStream(struct LINNEEncoder *encoder, void *cuda_stream);
void LINNEB200_DecoderUseStream(struct LINNEDecoder *decoder, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
