/* linne_tables.h -- format constants of the .lnn bitstream shared by host C and CUDA code.
 *
 * These are properties of the FORMAT (any conforming implementation carries them):
 *   presets        reference libs/linne_internal/src/linne_internal.c:16-41
 *   coefficient symbol statistics (fixed Huffman model)   linne_internal.c:26-28
 *   bit widths / thresholds                               libs/linne_internal/include/linne_internal.h:8-35
 */
#ifndef LINNE_B200_TABLES_H
#define LINNE_B200_TABLES_H

#include <stdint.h>

#define LNB_MAX_CHANNELS       8
#define LNB_MAX_LAYERS         3
#define LNB_MAX_PARAMS         128
#define LNB_MAX_LAMBDAS        4
#define LNB_MAX_LEVELS         8      /* unit counts 1,2,4,...,128 */
#define LNB_MAX_UNITS          128
#define LNB_HEADER_SIZE        30
#define LNB_BLOCK_HEADER_SIZE  11
#define LNB_SYNC_CODE          0xFFFFu
#define LNB_PREEM_SHIFT        5
#define LNB_NUM_PREEM          2
#define LNB_MAX_PORDER         10
#define LNB_MAX_BITS_PER_SAMPLE 31    /* the (bits + 1)-bit pre-emphasis state must fit a 32-bit field (bit_stream.h:317) */
#define LNB_MAX_PARTITIONS     1024
#define LNB_RAW_THRESHOLD      0.95f  /* float on purpose: linne_internal.h:24 */

enum { LNB_BLOCK_COMPRESSED = 0, LNB_BLOCK_SILENT = 1, LNB_BLOCK_RAW = 2 };

typedef struct LnbPreset {
    int32_t num_layers;
    int32_t layer_params[LNB_MAX_LAYERS];
    int32_t num_lambdas;
    double  lambdas[LNB_MAX_LAMBDAS];
} LnbPreset;

#ifdef __cplusplus
extern "C" {
#endif
extern const LnbPreset g_lnb_presets[8];
extern const uint32_t  g_lnb_coef_freq[256];
#ifdef __cplusplus
}
#endif

#endif
