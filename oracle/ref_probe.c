/* oracle/ref_probe.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Builds into oracle/_ref/liblinne_ref.so together with the unmodified reference sources
 * (see oracle/Makefile).  It textually includes the reference encoder/decoder translation
 * units -- exactly what the reference's own white-box tests do
 * (test/linne_encoder/linne_encoder_test.cpp:7-9, test/linne_decoder/linne_decoder_test.cpp:10)
 * -- so that the tests can read the per-block analysis results (unit counts, right shifts,
 * quantised coefficients, pre-emphasis state) which the reference keeps inside its opaque
 * handle.  No reference source is copied into this repository.
 */
#include "linne_encoder.c"
#include "linne_decoder.c"

/* Results of the most recent EncodeBlock on this handle (linne_encoder.c:17-42 members). */
void RefProbe_EncoderLastParams(const struct LINNEEncoder *enc, uint32_t ch, uint32_t layer,
        uint32_t *num_units, uint32_t *rshift, int32_t *coef_out, uint32_t num_coef)
{
    uint32_t i;
    *num_units = enc->num_units[ch][layer];
    *rshift = enc->rshifts[ch][layer];
    for (i = 0; i < num_coef; i++) coef_out[i] = enc->params_int[ch][layer][i];
}

void RefProbe_EncoderLastPreemphasis(const struct LINNEEncoder *enc, uint32_t ch,
        int32_t prev_out[2], int32_t coef_out[2])
{
    uint32_t l;
    for (l = 0; l < LINNE_NUM_PREEMPHASIS_FILTERS; l++) {
        prev_out[l] = enc->pre_emphasis_prev[ch][l];
        coef_out[l] = enc->pre_emphasis[ch][l].coef;
    }
}

/* Residual of the last compressed block, channel ch (valid for num_samples of that block). */
const int32_t *RefProbe_EncoderLastResidual(const struct LINNEEncoder *enc, uint32_t ch)
{
    return enc->residual[ch];
}

/* Direct handles on a few internal routines for known-answer checks of the oracle. */
uint16_t RefProbe_CRC16(const uint8_t *data, uint64_t size)
{
    return LINNEUtility_CalculateCRC16(data, size);
}

void RefProbe_HuffmanCodes(const uint32_t *counts, uint32_t n, uint32_t *code_out, uint8_t *len_out)
{
    struct StaticHuffmanTree tree;
    struct StaticHuffmanCodes codes;
    uint32_t i;
    StaticHuffman_BuildHuffmanTree(counts, n, &tree);
    StaticHuffman_ConvertTreeToCodes(&tree, &codes);
    for (i = 0; i < n; i++) { code_out[i] = codes.codes[i].code; len_out[i] = codes.codes[i].bit_count; }
}

const uint32_t *RefProbe_CoefFreqTable(uint32_t preset)
{
    return g_linne_parameter_preset[preset].coef_symbol_freq_table;
}

void RefProbe_QuantizeCoefficients(const double *coef, uint32_t order, int32_t *int_coef, uint32_t *rshift)
{
    LPC_QuantizeCoefficients(coef, order, LINNE_LPC_COEFFICIENT_BITWIDTH, int_coef, rshift);
}
