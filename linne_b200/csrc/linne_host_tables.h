/* linne_host_tables.h -- tables computed on the host at start-up (see linne_tables.c). */
#ifndef LINNE_B200_HOST_TABLES_H
#define LINNE_B200_HOST_TABLES_H

#include <stdint.h>

#define LNB_HUFF_LUT_BITS       14
#define LNB_NUM_K2_THRESHOLDS   40

typedef struct LnbHostTables {
    uint32_t huff_code[256];
    uint8_t  huff_len[256];
    uint32_t huff_max_len;
    uint16_t huff_lut[1u << LNB_HUFF_LUT_BITS];      /* (symbol << 4) | code length */
    double   k2_threshold[LNB_NUM_K2_THRESHOLDS];    /* k2(mean) = #{k >= 1 : mean >= threshold[k]} */
    uint16_t crc_table[256];
} LnbHostTables;

#ifdef __cplusplus
extern "C" {
#endif
const LnbHostTables *lnb_tables_get(void);
double lnb_welch_scale(uint32_t unit_len);
#ifdef __cplusplus
}
#endif

#endif
