/* linne.h -- shared constants, result codes and the stream header of the LINNE lossless codec.
 *
 * DROP-IN BOUNDARY.  This header reproduces, declaration for declaration, the ABI contract of the
 * reference's include/linne.h (constants :7-19, LINNEApiResult :22-31, LINNEChannelProcessMethod
 * :34-38, struct LINNEHeader :41-51).  Enumerator values, struct member order and member types
 * are identical, so code compiled against the reference header links and runs against
 * liblinne_b200.so unchanged.  Behaviour behind these declarations is implemented from scratch
 * as CUDA kernels for sm_100a.
 */
#ifndef LINNE_H_INCLUDED
#define LINNE_H_INCLUDED

#include "linne_stdint.h"

/* .lnn container revision written into / required from the header (reference linne.h:7) */
#define LINNE_FORMAT_VERSION        1
/* bitstream revision of the block payload (reference linne.h:10) */
#define LINNE_CODEC_VERSION         2
/* serialized size of struct LINNEHeader in bytes (reference linne.h:13) */
#define LINNE_HEADER_SIZE           30
/* channel limit of one stream (reference linne.h:16) */
#define LINNE_MAX_NUM_CHANNELS      8
/* number of -m presets, 0..7 (reference linne.h:19) */
#define LINNE_NUM_PARAMETER_PRESETS 8

/* Result code of every API call (reference linne.h:22-31; values 0..7 in this order). */
typedef enum LINNEApiResultTag {
    LINNE_APIRESULT_OK = 0,
    LINNE_APIRESULT_INVALID_ARGUMENT,
    LINNE_APIRESULT_INVALID_FORMAT,
    LINNE_APIRESULT_INSUFFICIENT_BUFFER,
    LINNE_APIRESULT_INSUFFICIENT_DATA,
    LINNE_APIRESULT_PARAMETER_NOT_SET,
    LINNE_APIRESULT_DETECT_DATA_CORRUPTION,
    LINNE_APIRESULT_NG
} LINNEApiResult;

/* Inter-channel decorrelation applied to channels 0/1 (reference linne.h:34-38). */
typedef enum LINNEChannelProcessMethodTag {
    LINNE_CH_PROCESS_METHOD_NONE = 0,
    LINNE_CH_PROCESS_METHOD_MS,
    LINNE_CH_PROCESS_METHOD_INVALID
} LINNEChannelProcessMethod;

/* Stream header, host representation (reference linne.h:41-51). */
struct LINNEHeader {
    uint32_t format_version;
    uint32_t codec_version;
    uint16_t num_channels;
    uint32_t num_samples;              /* per channel, whole stream */
    uint32_t sampling_rate;
    uint16_t bits_per_sample;
    uint32_t num_samples_per_block;
    uint8_t preset;
    LINNEChannelProcessMethod ch_process_method;
};

#endif /* LINNE_H_INCLUDED */
