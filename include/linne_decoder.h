/* linne_decoder.h -- decoder half of the LINNE C API, served by the B200 implementation.
 *
 * DROP-IN BOUNDARY.  Same seven entry points, argument meaning and error ordering as the
 * reference's include/linne_decoder.h:23-50 (implemented in the reference by
 * libs/linne_decoder/src/linne_decoder.c); struct layout equals linne_decoder.h:8-13.
 * If no usable CUDA device/driver is present, Create returns NULL (there is NO CPU fallback).
 */
#ifndef LINNE_DECODER_H_INCLUDED
#define LINNE_DECODER_H_INCLUDED

#include "linne.h"
#include "linne_stdint.h"

/* Capacity / behaviour of a handle (reference linne_decoder.h:8-13). */
struct LINNEDecoderConfig {
    uint32_t max_num_channels;
    uint32_t max_num_layers;
    uint32_t max_num_parameters_per_layer;
    uint8_t check_crc;                  /* 1: verify the per-block CRC16, anything else: skip */
};

struct LINNEDecoder;

#ifdef __cplusplus
extern "C" {
#endif

/* Parse a 30-byte stream header.  Replaces linne_decoder.c:60-131. */
LINNEApiResult LINNEDecoder_DecodeHeader(
        const uint8_t *data, uint32_t data_size, struct LINNEHeader *header);

/* Bytes of host work area Create needs, or -1.  Replaces linne_decoder.c:187-216. */
int32_t LINNEDecoder_CalculateWorkSize(const struct LINNEDecoderConfig *config);

/* Build a handle.  Replaces linne_decoder.c:219-296. */
struct LINNEDecoder *LINNEDecoder_Create(
        const struct LINNEDecoderConfig *config, void *work, int32_t work_size);

/* Release the handle.  Replaces linne_decoder.c:299-306. */
void LINNEDecoder_Destroy(struct LINNEDecoder *decoder);

/* Validate and latch a stream header.  Replaces linne_decoder.c:309-354. */
LINNEApiResult LINNEDecoder_SetHeader(
        struct LINNEDecoder *decoder, const struct LINNEHeader *header);

/* Decode one block (batch of one).  Replaces linne_decoder.c:564-668. */
LINNEApiResult LINNEDecoder_DecodeBlock(
        struct LINNEDecoder *decoder,
        const uint8_t *data, uint32_t data_size,
        int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples,
        uint32_t *decode_size, uint32_t *num_decode_samples);

/* Decode header + every block of a stream in one batched GPU pass.  Replaces linne_decoder.c:671-730. */
LINNEApiResult LINNEDecoder_DecodeWhole(
        struct LINNEDecoder *decoder,
        const uint8_t *data, uint32_t data_size,
        int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples);

#ifdef __cplusplus
}
#endif

#endif /* LINNE_DECODER_H_INCLUDED */
