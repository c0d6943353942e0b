/* linne_encoder.h -- encoder half of the LINNE C API, served by the B200 implementation.
 *
 * DROP-IN BOUNDARY.  Same seven entry points, same argument meaning and the same error
 * behaviour as the reference's include/linne_encoder.h:35-61 (implemented in the reference by
 * libs/linne_encoder/src/linne_encoder.c); struct layouts equal linne_encoder.h:8-25.
 *
 * Differences a caller can observe (all documented in DESIGN.md / INTEGRATION.md):
 *  - the work area given to Create only holds the host-side handle; device memory lives inside
 *    the handle and is sized on first use (grow-only);
 *  - EncodeWhole batches every block x channel of the call into one set of kernel launches;
 *  - the compressed path returns LINNE_APIRESULT_INSUFFICIENT_BUFFER when the output buffer is
 *    too small (the reference overruns silently in Release, linne_encoder.c:699-749);
 *  - if no usable CUDA device/driver is present, Create returns NULL (there is NO CPU fallback).
 */
#ifndef LINNE_ENCODER_H_INCLUDED
#define LINNE_ENCODER_H_INCLUDED

#include "linne.h"
#include "linne_stdint.h"

/* Per-stream encode parameters (reference linne_encoder.h:8-17). */
struct LINNEEncodeParameter {
    uint16_t num_channels;
    uint16_t bits_per_sample;
    uint32_t sampling_rate;
    uint16_t num_samples_per_block;
    uint8_t preset;                               /* 0..7, the CLI's -m */
    LINNEChannelProcessMethod ch_process_method;
    uint8_t enable_learning;                      /* CLI -l : momentum-SGD refinement */
    uint8_t num_afmethod_iterations;              /* CLI -a : IRLS iterations, 0 = off */
};

/* Capacity of a handle (reference linne_encoder.h:20-25). */
struct LINNEEncoderConfig {
    uint32_t max_num_channels;
    uint32_t max_num_samples_per_block;
    uint32_t max_num_layers;
    uint32_t max_num_parameters_per_layer;
};

struct LINNEEncoder;

#ifdef __cplusplus
extern "C" {
#endif

/* Serialize a 30-byte stream header.  Replaces linne_encoder.c:53-138. */
LINNEApiResult LINNEEncoder_EncodeHeader(
        const struct LINNEHeader *header, uint8_t *data, uint32_t data_size);

/* Bytes of host work area Create needs, or -1 for a bad config.  Replaces linne_encoder.c:201-265. */
int32_t LINNEEncoder_CalculateWorkSize(const struct LINNEEncoderConfig *config);

/* Build a handle in `work` (or self-allocate when work==NULL && work_size==0).
 * Replaces linne_encoder.c:268-396. */
struct LINNEEncoder *LINNEEncoder_Create(
        const struct LINNEEncoderConfig *config, void *work, int32_t work_size);

/* Release device resources (and the work area if self-allocated).  Replaces linne_encoder.c:399-407. */
void LINNEEncoder_Destroy(struct LINNEEncoder *encoder);

/* Validate and latch stream parameters.  Replaces linne_encoder.c:410-477. */
LINNEApiResult LINNEEncoder_SetEncodeParameter(
        struct LINNEEncoder *encoder, const struct LINNEEncodeParameter *parameter);

/* Encode one block (batch of one).  Replaces linne_encoder.c:774-862. */
LINNEApiResult LINNEEncoder_EncodeBlock(
        struct LINNEEncoder *encoder,
        const int32_t *const *input, uint32_t num_samples,
        uint8_t *data, uint32_t data_size, uint32_t *output_size);

/* Encode header + every block of a stream in one batched GPU pass.  Replaces linne_encoder.c:865-932. */
LINNEApiResult LINNEEncoder_EncodeWhole(
        struct LINNEEncoder *encoder,
        const int32_t *const *input, uint32_t num_samples,
        uint8_t *data, uint32_t data_size, uint32_t *output_size);

#ifdef __cplusplus
}
#endif

#endif /* LINNE_ENCODER_H_INCLUDED */
