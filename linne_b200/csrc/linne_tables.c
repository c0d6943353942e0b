/* linne_tables.c -- host-side construction of the constant tables the kernels consume.
 *
 * Built once per process (lnb_tables_get) and uploaded to the device by the CUDA shim:
 *   - coefficient Huffman code book + 14-bit decode look-up table
 *       (tree construction rule of reference libs/static_huffman/src/static_huffman.c:28-131:
 *        merge the two smallest live nodes, index order breaks ties, first minimum is child 0)
 *   - recursive-Rice parameter thresholds: k2 is a monotone step function of the partition mean
 *       (reference libs/linne_coder/src/linne_coder.c:172-200); the step positions are found on
 *       the host with the reference's libm formula so that the device reproduces k2 exactly by
 *       comparison, without device-side log()
 *   - CRC16-IBM byte table (reference libs/linne_internal/src/linne_utility.c:8-41 holds the same
 *       256 values as literals; here they are generated from the polynomial 0xA001)
 */
#include "linne_tables.h"
#include "linne_host_tables.h"

#include <math.h>
#include <string.h>

const LnbPreset g_lnb_presets[8] = {
    { 2, { 2,  32,  0 }, 1, { 0.0, 0.0, 0.0, 0.0 } },
    { 2, { 2,  32,  0 }, 2, { 0.0, 1.0 / 512.0, 0.0, 0.0 } },
    { 3, { 4,  64,  8 }, 1, { 0.0, 0.0, 0.0, 0.0 } },
    { 3, { 4,  64,  8 }, 2, { 0.0, 1.0 / 512.0, 0.0, 0.0 } },
    { 3, { 4,  64,  8 }, 4, { 0.0, 1.0 / 2048.0, 1.0 / 512.0, 1.0 / 128.0 } },
    { 3, { 4, 128, 16 }, 1, { 0.0, 0.0, 0.0, 0.0 } },
    { 3, { 4, 128, 16 }, 2, { 0.0, 1.0 / 512.0, 0.0, 0.0 } },
    { 3, { 4, 128, 16 }, 4, { 0.0, 1.0 / 2048.0, 1.0 / 512.0, 1.0 / 128.0 } },
};

const uint32_t g_lnb_coef_freq[256] = {
    2944693,2417040,2500224,2220717,2361506,2005548,2161319,1804396,1961813,1628891,1774159,1471673,
    1604885,1335449,1451476,1218111,1316402,1112581,1200154,1019661,1094294,935533,1000598,861453,
    914647,793863,837607,733372,769686,679634,709504,630828,653277,583990,602876,545068,556612,507071,
    516014,473301,478009,441389,442848,415057,412045,389010,384623,364872,359578,343600,335976,322541,
    314173,304513,293388,286871,277191,271905,260699,256892,245269,243815,231142,231894,217938,220197,
    205798,209146,196061,199652,185811,189659,176121,181265,168122,173827,159699,167156,150968,158868,
    144276,152666,137117,146329,130245,141026,124044,134984,118946,130389,113141,125287,108826,120399,
    102664,116857,98953,112210,93718,109059,89757,106036,86363,102597,82554,99558,78306,96473,76105,
    92575,72428,89227,68911,85952,66258,82764,63571,80241,61196,78050,58502,75544,56329,73454,53557,
    71750,51667,81769,52853,90325,53934,86990,51338,83565,48756,80882,47304,78156,44823,75050,43129,
    72304,41339,70163,39767,67853,37538,65134,35572,62994,34367,61059,32981,58664,31690,56196,30505,
    54354,29091,52803,27750,50577,26523,49428,25414,47359,24109,46224,23419,44925,22167,43578,21336,
    42201,20551,41434,19640,39842,18815,38775,18200,37804,17159,36516,16591,35217,16053,34221,14962,
    33101,14533,32077,13842,31550,13427,30277,12962,29616,12296,29090,11678,27922,11467,27212,10733,
    26329,10270,25938,9930,24828,9336,24672,9085,23868,8616,23456,8430,22633,7892,21759,7594,21723,
    7430,20729,6988,20475,6673,20100,6489,19480,6100,18993,5912,18480,5599,17993,5292,17267,5100,
    17013,4919,16502,4721,16304,4471,16040,4313,16120,4090,17146,3921,28239,3817,49638,5544,7587,
};

/* ---------------------------------------------------------------------------------------------- */
static void build_huffman(const uint32_t *freq, uint32_t n, uint32_t *codes, uint8_t *lens)
{
    uint32_t weight[512], kid0[512], kid1[512];
    uint32_t alive = n, top = n, i;
    struct { uint32_t node, code; uint8_t len; } todo[512];
    int depth = 0;

    memset(weight, 0, sizeof(weight));
    for (i = 0; i < n; i++) weight[i] = freq[i] ? freq[i] : 1u;
    while (alive > 1) {
        uint32_t lo = 0xFFFFFFFFu, lo2 = 0xFFFFFFFFu;     /* indices of smallest / second smallest */
        for (i = 0; i < top; i++) {
            if (!weight[i]) continue;
            if (lo == 0xFFFFFFFFu || weight[i] < weight[lo]) { lo2 = lo; lo = i; }
            else if (lo2 == 0xFFFFFFFFu || weight[i] < weight[lo2]) { lo2 = i; }
        }
        weight[top] = weight[lo] + weight[lo2];
        weight[lo] = weight[lo2] = 0;
        kid0[top] = lo; kid1[top] = lo2;
        top++; alive--;
    }
    todo[depth].node = top - 1; todo[depth].code = 0; todo[depth].len = 0; depth++;
    while (depth) {
        uint32_t node, code; uint8_t len;
        depth--;
        node = todo[depth].node; code = todo[depth].code; len = todo[depth].len;
        if (node < n) { codes[node] = code; lens[node] = len; continue; }
        todo[depth].node = kid1[node]; todo[depth].code = (code << 1) | 1u; todo[depth].len = (uint8_t)(len + 1); depth++;
        todo[depth].node = kid0[node]; todo[depth].code = code << 1;        todo[depth].len = (uint8_t)(len + 1); depth++;
    }
}

/* k2 for a partition mean, with the reference's formula and the host libm (linne_coder.c:180-184) */
static uint32_t rice_k2_formula(double mean)
{
    const double optx = 0.5127629514437670454896078808815218508243560791015625;
    const double rho = 1.0 / (1.0 + mean);
    const double f = floor(log(log(optx) / log(1.0 - rho)) * 1.4426950408889634);
    return (uint32_t)((0 > f) ? 0 : f);
}

/* smallest non-negative double m with rice_k2_formula(m) >= k  (bisection on the bit pattern) */
static double k2_threshold(uint32_t k)
{
    union { double d; uint64_t u; } lo, hi, mid;
    lo.d = 0.0; hi.d = 1.0e12;
    if (rice_k2_formula(hi.d) < k) return INFINITY;
    while (hi.u - lo.u > 1) {
        mid.u = lo.u + (hi.u - lo.u) / 2;
        if (rice_k2_formula(mid.d) >= k) hi = mid; else lo = mid;
    }
    return hi.d;
}

static LnbHostTables g_tables;
static int g_tables_ready = 0;

const LnbHostTables *lnb_tables_get(void)
{
    uint32_t sym, i, k;
    if (g_tables_ready) return &g_tables;

    build_huffman(g_lnb_coef_freq, 256, g_tables.huff_code, g_tables.huff_len);
    g_tables.huff_max_len = 0;
    for (sym = 0; sym < 256; sym++)
        if (g_tables.huff_len[sym] > g_tables.huff_max_len) g_tables.huff_max_len = g_tables.huff_len[sym];
    /* decode LUT indexed by the next LNB_HUFF_LUT_BITS bits of the stream: (symbol << 4) | length */
    memset(g_tables.huff_lut, 0, sizeof(g_tables.huff_lut));
    for (sym = 0; sym < 256; sym++) {
        const uint32_t len = g_tables.huff_len[sym];
        const uint32_t span = 1u << (LNB_HUFF_LUT_BITS - len);
        const uint32_t first = g_tables.huff_code[sym] << (LNB_HUFF_LUT_BITS - len);
        for (i = 0; i < span; i++) g_tables.huff_lut[first + i] = (uint16_t)((sym << 4) | len);
    }

    g_tables.k2_threshold[0] = 0.0;
    for (k = 1; k < LNB_NUM_K2_THRESHOLDS; k++) g_tables.k2_threshold[k] = k2_threshold(k);

    for (i = 0; i < 256; i++) {
        uint16_t c = (uint16_t)i;
        for (k = 0; k < 8; k++) c = (uint16_t)((c & 1) ? (c >> 1) ^ 0xA001 : c >> 1);
        g_tables.crc_table[i] = c;
    }
    g_tables_ready = 1;
    return &g_tables;
}

double lnb_welch_scale(uint32_t unit_len)
{
    /* 4 / (m-1)^2 through pow(), exactly as the reference's window does (libs/lpc/src/lpc.c:199):
     * pow() and a plain division differ in the last bit for a few lengths, so the host computes it. */
    return 4.0 * pow((double)(unit_len - 1u), -2.0);
}
