/* linne_oracle.h -- CPU restatement of the LINNE block encode/decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed from
 * the product library (linne_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liblinne_oracle.so, and only as the checker /
 * reported CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement against the
 * unmodified reference built by oracle/Makefile (`make ref` -> oracle/_ref/liblinne_ref.so):
 * byte-identical streams at every preset on synthetic clips and all test generators, plus the
 * reference's own KATs (CRC16: test/linne_internal/main.cpp:25-32; Huffman:
 * test/static_huffman/main.cpp:40-73) and the committed golden streams under tests/golden/.
 *
 * Every function cites the reference file:line whose behaviour it restates (paths relative to
 * the reference root).  The code is written from the algorithm, stage by stage, in the
 * decomposition the CUDA kernels use, so stage outputs can be compared one to one.
 */
#ifndef LINNE_ORACLE_H_INCLUDED
#define LINNE_ORACLE_H_INCLUDED

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LO_MAX_CHANNELS   8
#define LO_MAX_LAYERS     3
#define LO_MAX_PARAMS     128
#define LO_MAX_LAMBDAS    4
#define LO_HEADER_SIZE    30
#define LO_BLOCK_HEADER   11

/* result codes: numerically equal to LINNEApiResult (include/linne.h:22-31) */
enum { LO_OK = 0, LO_INVALID_ARGUMENT, LO_INVALID_FORMAT, LO_INSUFFICIENT_BUFFER,
       LO_INSUFFICIENT_DATA, LO_PARAMETER_NOT_SET, LO_DATA_CORRUPTION, LO_NG };

/* block payload kinds (libs/linne_internal/include/linne_internal.h:50-55) */
enum { LO_BLOCK_COMPRESSED = 0, LO_BLOCK_SILENT = 1, LO_BLOCK_RAW = 2 };

typedef struct lo_preset {
    int num_layers;
    int layer_params[LO_MAX_LAYERS];
    int num_lambdas;
    double lambdas[LO_MAX_LAMBDAS];
} lo_preset;

typedef struct lo_stream_info {          /* fields of the 30-byte header */
    uint32_t format_version, codec_version;
    uint32_t num_channels, num_samples, sampling_rate, bits_per_sample, block_size;
    uint32_t preset, ms;
} lo_stream_info;

/* Per block-channel record of every intermediate the encoder produced for the most recent
 * compressed block (what the stage-level parity tests compare against). */
typedef struct lo_channel_trace {
    int32_t  preem_prev[2], preem_coef[2];
    uint32_t num_units[LO_MAX_LAYERS];
    uint32_t rshift[LO_MAX_LAYERS];
    int32_t  coef[LO_MAX_LAYERS][LO_MAX_PARAMS];
    double   coef_f64[LO_MAX_LAYERS][LO_MAX_PARAMS];
    uint32_t porder;                      /* chosen partition order of the residual coder */
    uint32_t residual_bits;               /* bits the residual coder emitted for this channel */
} lo_channel_trace;

typedef struct lo_encoder lo_encoder;

/* ---- tables and small pure functions ---- */
const lo_preset *lo_get_preset(uint32_t preset);
uint16_t lo_crc16(const uint8_t *data, size_t size);
void lo_huffman_build(const uint32_t *counts, uint32_t n, uint32_t *codes, uint8_t *lengths);
void lo_coef_huffman_table(uint32_t codes[256], uint8_t lengths[256]);
const uint32_t *lo_coef_freq_table(void);
void lo_rice_parameter(double mean, uint32_t *k1, uint32_t *k2);
uint32_t lo_rice_code_length(uint32_t k1, uint32_t k2, uint32_t uval);

/* ---- header ---- */
int lo_header_encode(const lo_stream_info *h, uint8_t *out, uint32_t cap);
int lo_header_decode(const uint8_t *data, uint32_t size, lo_stream_info *h);

/* ---- integer signal stages (all in place unless noted) ---- */
void lo_ms_forward(int32_t *ch0, int32_t *ch1, uint32_t n);
void lo_ms_inverse(int32_t *ch0, int32_t *ch1, uint32_t n);
int32_t lo_preemphasis_coef(const int32_t *x, uint32_t n);
void lo_preemphasis(int32_t *x, uint32_t n, int32_t prev, int32_t coef);
void lo_deemphasis2(int32_t *x, uint32_t n, const int32_t prev[2], const int32_t coef[2]);
void lo_predict(const int32_t *in, uint32_t n, const int32_t *coef, uint32_t num_params,
                uint32_t rshift, uint32_t num_units, int32_t *residual);
void lo_synthesize(int32_t *data, uint32_t n, const int32_t *coef, uint32_t num_params,
                   uint32_t rshift, uint32_t num_units);
void lo_quantize(const double *coef, uint32_t n, int32_t *q, uint32_t *rshift);

/* ---- residual coder ---- */
/* choose partition order + per-partition k2; returns total bits (incl. 10-bit porder field) */
uint32_t lo_coder_plan(const int32_t *res, uint32_t n, uint32_t *porder, uint32_t *k2_out /*[1024]*/);

/* ---- encoder / decoder objects ---- */
lo_encoder *lo_encoder_create(uint32_t max_channels, uint32_t max_block);
void lo_encoder_destroy(lo_encoder *e);
int lo_encoder_configure(lo_encoder *e, uint32_t num_channels, uint32_t bits, uint32_t rate,
                         uint32_t block_size, uint32_t preset, uint32_t ms,
                         uint32_t enable_learning, uint32_t af_iterations);
int lo_encode_block(lo_encoder *e, const int32_t *const *input, uint32_t n,
                    uint8_t *out, uint32_t cap, uint32_t *out_size);
int lo_encode_whole(lo_encoder *e, const int32_t *const *input, uint32_t num_samples,
                    uint8_t *out, uint32_t cap, uint32_t *out_size);
const lo_channel_trace *lo_encoder_trace(const lo_encoder *e, uint32_t ch);
const int32_t *lo_encoder_last_residual(const lo_encoder *e, uint32_t ch);
int lo_encoder_last_block_type(const lo_encoder *e);

/* Encode one block with FORCED analysis results (units / rshift / coefs and pre-emphasis taken
 * from `forced[ch]`): the "given identical quantised coefficients, bytes must be identical" leg. */
int lo_encode_block_forced(lo_encoder *e, const int32_t *const *input, uint32_t n,
                           const lo_channel_trace *forced, uint8_t *out, uint32_t cap, uint32_t *out_size);

/* analysis only: double samples (already normalised) -> units + double params per layer */
void lo_analyze(lo_encoder *e, const double *x, uint32_t n_analyze,
                uint32_t num_units[LO_MAX_LAYERS], double params[LO_MAX_LAYERS][LO_MAX_PARAMS]);

int lo_decode_block(const lo_stream_info *h, int check_crc, const uint8_t *data, uint32_t size,
                    int32_t **out, uint32_t out_channels, uint32_t out_samples,
                    uint32_t *consumed, uint32_t *decoded_samples);
int lo_decode_whole(const uint8_t *data, uint32_t size, int check_crc,
                    int32_t **out, uint32_t out_channels, uint32_t out_samples);

#ifdef __cplusplus
}
#endif
#endif
