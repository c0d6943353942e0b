/* linne_decoder_host.c -- LINNEDecoder_* entry points (include/linne_decoder.h) on top of the CUDA shim.
 *
 * Replaces reference libs/linne_decoder/src/linne_decoder.c.  The host side does what is serial and
 * tiny (argument checks, the 30-byte header, hopping over the block size fields to build the block
 * table); everything that touches samples or payload bits runs on the GPU, batched over every block
 * of the call.  Error results and their precedence follow the reference (cited per check).
 */
#define _POSIX_C_SOURCE 200112L
#include "linne_decoder.h"
#include "linne_b200.h"
#include "lnb_host_util.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

/* LINNE_B200_TRACE=1: host-side time stamps of a range's phases on stderr (debugging aid) */
static int dec_trace_on(void) { static int on = -1; if (on < 0) { const char *e = getenv("LINNE_B200_TRACE"); on = (e && *e == '1') ? 1 : 0; } return on; }
static double dec_now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
#define DEC_TRACE(dec, what) do { if (dec_trace_on()) fprintf(stderr, "[trace] range %2u %-16s %.3f\n", (unsigned)(dec)->turn_index, what, dec_now_ms()); } while (0)

#define DEC_FLAG_OWN_WORK   (1u << 0)
#define DEC_FLAG_HEADER_SET (1u << 1)
#define DEC_FLAG_CHECK_CRC  (1u << 2)

struct LINNEDecoder {
    struct LINNEHeader header;
    uint32_t max_num_channels, max_num_layers, max_num_parameters_per_layer;
    uint32_t flags;
    void *work;
    LnbDevice *dev;
    LnbBuf d_stream, d_blocks, d_params, d_pcm, d_packed;
    LnbBuf h_blocks;                       /* pinned LnbBlockDesc[] */
    LnbBuf h_stream;                       /* pinned copy of a device-resident stream (block hop on the host: fallback only) */
    LnbBuf d_hop_files, d_hop_res, d_hop_table, h_hop;     /* device-side block hop: records, results, table; pinned staging */
    LnbBuf h_pcm;                          /* pinned int32 [C][stage_stride]: PCM staging of DecodeBlock / read-ahead cache */
    /* DecodeBlock read-ahead: blocks decoded ahead of the caller in one batch and served from here */
    uint32_t readahead;                    /* blocks decoded per batch by DecodeBlock (<= 1: batch of one) */
    uint32_t tput_min_blocks;              /* batches of at least this many blocks take the throughput kernels (0 = never) */
    uint32_t stage_stride;                 /* samples per plane of h_pcm in the last staged call */
    uint32_t ra_count, ra_next;            /* cached blocks / next one to serve */
    uint8_t *ra_image;                     /* the cached blocks' bytes (validated against the caller's on every hit) */
    size_t ra_image_cap;
    struct LnbCachedBlock { uint32_t byte_off, byte_size, smp_off, nsmp; } *ra_blocks;
    uint32_t ra_blocks_cap;
    /* several GPUs behind one handle: DecodeWhole splits the block table into contiguous ranges, one child
     * handle (own device, own host thread) per range */
    uint32_t num_devices;
    struct LINNEDecoder *child[LNB_MAX_DEVICES];
    LnbTurnstile *turn;                    /* a child: the turnstile of the call it works for (else NULL) and its range index */
    uint32_t turn_index;
    int rank_override;                     /* >= 0: the stream priority rank of a child (its place in the pipeline) instead of the preset's */
    struct LINNEDecoderConfig config;
    LnbBlockDesc *shard_table;             /* host block table of the sharding hop */
    uint32_t shard_table_cap;
};
#define LNB_MAX_READAHEAD 1024u
#define LNB_READAHEAD_MAX_SAMPLES (1u << 21)  /* per channel and batch: bounds the pinned PCM cache (8 ch: 64 MiB) */
#define LNB_TPUT_MIN_BLOCKS 2560ul

/* reference linne_decoder.c:60-131 */
LINNEApiResult LINNEDecoder_DecodeHeader(const uint8_t *data, uint32_t data_size, struct LINNEHeader *header)
{
    if (data == NULL || header == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if (data_size < LINNE_HEADER_SIZE) return LINNE_APIRESULT_INSUFFICIENT_DATA;
    if (data[0] != 'I' || data[1] != 'B' || data[2] != 'R' || data[3] != 'A') return LINNE_APIRESULT_INVALID_FORMAT;
    lnb_header_read(data, header);
    return LINNE_APIRESULT_OK;
}

static int config_ok(const struct LINNEDecoderConfig *c)
{
    return c && c->max_num_channels && c->max_num_layers && c->max_num_parameters_per_layer;
}

/* reference linne_decoder.c:187-216: the work area only has to hold the host handle here */
int32_t LINNEDecoder_CalculateWorkSize(const struct LINNEDecoderConfig *config)
{
    if (!config_ok(config)) return -1;
    return (int32_t)(sizeof(struct LINNEDecoder) + LNB_ALIGNMENT);
}

/* reference linne_decoder.c:219-296 */
struct LINNEDecoder *LINNEDecoder_Create(const struct LINNEDecoderConfig *config, void *work, int32_t work_size)
{
    struct LINNEDecoder *dec;
    int own = 0;
    if (work == NULL && work_size == 0) {
        if ((work_size = LINNEDecoder_CalculateWorkSize(config)) < 0) return NULL;
        work = malloc((size_t)work_size);
        own = 1;
    }
    if (!config_ok(config) || work == NULL || work_size < LINNEDecoder_CalculateWorkSize(config)) {
        if (own) free(work);
        return NULL;
    }
    dec = (struct LINNEDecoder *)LNB_ROUNDUP((uintptr_t)work, LNB_ALIGNMENT);
    memset(dec, 0, sizeof(*dec));
    dec->rank_override = -1;
    dec->work = work;
    dec->max_num_channels = config->max_num_channels;
    dec->max_num_layers = config->max_num_layers;
    dec->max_num_parameters_per_layer = config->max_num_parameters_per_layer;
    dec->config = *config;
    if (own) dec->flags |= DEC_FLAG_OWN_WORK;
    if (config->check_crc == 1) dec->flags |= DEC_FLAG_CHECK_CRC;
    {   /* LINNE_B200_READAHEAD=K: DecodeBlock decodes K blocks per batch (same results, see decode_block_cached) */
        const char *e = getenv("LINNE_B200_READAHEAD");
        const long k = e ? strtol(e, NULL, 10) : 0;
        dec->readahead = (k > 1) ? (uint32_t)(k > (long)LNB_MAX_READAHEAD ? (long)LNB_MAX_READAHEAD : k) : 0u;
    }
    {   /* LINNE_B200_TPUT_MIN_BLOCKS=N moves the switch-over point to the throughput decoder (0 = never) */
        const char *e = getenv("LINNE_B200_TPUT_MIN_BLOCKS");
        dec->tput_min_blocks = e ? (uint32_t)strtoul(e, NULL, 10) : (uint32_t)LNB_TPUT_MIN_BLOCKS;
    }
    if (lnb_shim_open(&dec->dev, -1) != 0) {
        fprintf(stderr, "linne_b200: no usable CUDA device -- the decoder has no CPU fallback\n");
        if (own) free(work);
        return NULL;
    }
    {   /* LINNE_B200_GPUS=N (or "all"): DecodeWhole of this handle shards its blocks over N devices */
        const char *e = getenv("LINNE_B200_GPUS");
        if (e && *e) LINNEB200_DecoderSetDevices(dec, (e[0] == 'a') ? (uint32_t)lnb_shim_device_count() : (uint32_t)strtoul(e, NULL, 10));
    }
    return dec;
}

/* reference linne_decoder.c:299-306 */
void LINNEDecoder_Destroy(struct LINNEDecoder *dec)
{
    uint32_t k;
    if (dec == NULL) return;
    for (k = 0; k < LNB_MAX_DEVICES; k++) if (dec->child[k]) { LINNEDecoder_Destroy(dec->child[k]); dec->child[k] = NULL; }
    free(dec->shard_table); dec->shard_table = NULL; dec->shard_table_cap = 0;
    if (dec->dev) {
        lnb_buf_release_device(dec->dev, &dec->d_stream);
        lnb_buf_release_device(dec->dev, &dec->d_blocks);
        lnb_buf_release_device(dec->dev, &dec->d_params);
        lnb_buf_release_device(dec->dev, &dec->d_pcm);
        lnb_buf_release_device(dec->dev, &dec->d_packed);
        lnb_buf_release_device(dec->dev, &dec->d_hop_files);
        lnb_buf_release_device(dec->dev, &dec->d_hop_res);
        lnb_buf_release_device(dec->dev, &dec->d_hop_table);
        lnb_buf_release_host(&dec->h_hop);
        lnb_buf_release_host(&dec->h_blocks);
        lnb_buf_release_host(&dec->h_stream);
        lnb_buf_release_host(&dec->h_pcm);
        free(dec->ra_image); dec->ra_image = NULL; dec->ra_image_cap = 0;
        free(dec->ra_blocks); dec->ra_blocks = NULL; dec->ra_blocks_cap = 0;
        dec->ra_count = dec->ra_next = 0;
        lnb_shim_close(dec->dev);
        dec->dev = NULL;
    }
    if (dec->flags & DEC_FLAG_OWN_WORK) free(dec->work);
}

/* reference linne_decoder.c:309-354 */
LINNEApiResult LINNEDecoder_SetHeader(struct LINNEDecoder *dec, const struct LINNEHeader *header)
{
    const LnbPreset *ps;
    int i;
    if (dec == NULL || header == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if (!lnb_header_fields_valid(header)) return LINNE_APIRESULT_INVALID_FORMAT;
    if (dec->max_num_channels < header->num_channels) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    ps = &g_lnb_presets[header->preset];
    if (dec->max_num_layers < (uint32_t)ps->num_layers) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    for (i = 0; i < ps->num_layers; i++)
        if (dec->max_num_parameters_per_layer < (uint32_t)ps->layer_params[i]) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    if (header->num_channels > LINNE_MAX_NUM_CHANNELS) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    dec->header = *header;
    dec->flags |= DEC_FLAG_HEADER_SET;
    dec->ra_count = dec->ra_next = 0;                          /* cached blocks belong to the previous header */
    lnb_shim_set_cost_rank(dec->dev, dec->rank_override >= 0 ? dec->rank_override : (int)header->preset);     /* longer predictors first (scheduling hint) */
    return LINNE_APIRESULT_OK;
}

/* ----------------------------------------------------------------------------------------------
 * Block table: hop over the size fields (serial by nature, a few ns per block).
 * Per block the reference checks, in this order (linne_decoder.c:604-653):
 *   sync (INVALID_FORMAT) -> size (INSUFFICIENT_DATA) -> CRC (DETECT_DATA_CORRUPTION)
 *   -> sample count vs buffer (INSUFFICIENT_BUFFER) -> type (INVALID_FORMAT, :651)
 *   -> raw payload size (INSUFFICIENT_DATA, :383)
 * Sync and size stop the hop; the later ones are recorded as the terminal block's `post_crc_error`
 * because a CRC mismatch (known only after the GPU pass) outranks them.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t num_blocks;          /* blocks entered in the table (the terminal block included if its CRC can be checked) */
    uint32_t num_decodable;       /* blocks that can be decoded (terminal block with a post-CRC error excluded) */
    LINNEApiResult framing_error; /* error that ends the stream after num_blocks blocks (or OK) */
    LINNEApiResult post_crc_error;/* error of block num_decodable (only when num_blocks == num_decodable + 1) */
    uint32_t total_samples;
    uint32_t end_offset;
} BlockScan;

static int scan_blocks(const struct LINNEHeader *h, const uint8_t *data, uint32_t data_size,
                       uint32_t start_offset, uint32_t max_samples, uint32_t sample_limit, uint32_t block_limit,
                       LnbBlockDesc *table, uint32_t table_cap, BlockScan *out)
{
    uint32_t off = start_offset, progress = 0, nb = 0;
    memset(out, 0, sizeof(*out));
    out->framing_error = LINNE_APIRESULT_OK;
    out->post_crc_error = LINNE_APIRESULT_OK;
    while (progress < sample_limit && off < data_size) {
        const uint8_t *p = data + off;
        const uint32_t remain = data_size - off;
        uint32_t size32, type, ns, consumed;
        LnbBlockDesc *d;
        if (nb >= table_cap) return 1;                         /* caller grows the table and retries */
        if (remain < 2 || lnb_rd_be(p, 2) != LNB_SYNC_CODE) {
            out->framing_error = (remain < 2) ? LINNE_APIRESULT_INSUFFICIENT_DATA : LINNE_APIRESULT_INVALID_FORMAT;
            break;
        }
        if (remain < 6) { out->framing_error = LINNE_APIRESULT_INSUFFICIENT_DATA; break; }
        size32 = lnb_rd_be(p + 2, 4);
        if ((uint64_t)size32 + 6u > remain) { out->framing_error = LINNE_APIRESULT_INSUFFICIENT_DATA; break; }
        if (size32 < 5u) { out->framing_error = LINNE_APIRESULT_INVALID_FORMAT; break; }   /* cannot hold crc+type+count */
        type = p[8];
        ns = lnb_rd_be(p + 9, 2);
        d = &table[nb];
        memset(d, 0, sizeof(*d));
        d->smp_off = progress; d->nsmp = ns; d->byte_off = off; d->byte_size = size32 + 6u; d->type = type;
        nb++;
        if (ns > max_samples - progress) { out->post_crc_error = LINNE_APIRESULT_INSUFFICIENT_BUFFER; break; }
        if (type > LNB_BLOCK_RAW) { out->post_crc_error = LINNE_APIRESULT_INVALID_FORMAT; break; }
        if (type == LNB_BLOCK_RAW) {
            const uint32_t need = (h->bits_per_sample * ns * h->num_channels) / 8u;
            if (remain - LNB_BLOCK_HEADER_SIZE < need) { out->post_crc_error = LINNE_APIRESULT_INSUFFICIENT_DATA; break; }
            consumed = LNB_BLOCK_HEADER_SIZE + (h->bits_per_sample / 8u) * ns * h->num_channels;
        } else if (type == LNB_BLOCK_SILENT) {
            consumed = LNB_BLOCK_HEADER_SIZE;
        } else {
            consumed = size32 + 6u;                            /* == 11 + ceil(bits/8) for any stream an encoder wrote */
        }
        out->num_decodable = nb;
        progress += ns;
        off += consumed;
        if (block_limit && nb >= block_limit) break;
    }
    out->num_blocks = nb;
    out->total_samples = progress;
    out->end_offset = off;
    return 0;
}

/* shared by DecodeBlock (block_limit >= 1, staged) and DecodeWhole (block_limit = 0) */
/* Resident mode (d_stream_ext / d_pcm_ext non-NULL): the stream image and/or the PCM planes already
 * live on the device, so the corresponding bulk copy is skipped; `data` (host) is still needed for
 * the block hop.
 * Staged mode (`staged`): the PCM comes back into the handle's pinned planes (dec->h_pcm, stride
 * dec->stage_stride) together with the block table, behind ONE synchronisation; the caller hands samples on
 * from there.  `scan_out` (optional) receives the hop's result; the block table stays in dec->h_blocks. */
static LINNEApiResult decode_range(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
                                   uint32_t start_offset, int32_t **buffer, uint32_t buffer_num_samples,
                                   uint32_t sample_limit, uint32_t block_limit, int staged,
                                   uint32_t *consumed_bytes, uint32_t *decoded_samples,
                                   const uint8_t *d_stream_ext, int32_t *d_pcm_ext, uint32_t pcm_stride_ext,
                                   BlockScan *scan_out, uint32_t *first_bad_out)
{
    const struct LINNEHeader *h = &dec->header;
    const uint32_t C = h->num_channels;
    LnbDecodeBatch batch;
    BlockScan scan;
    LnbBlockDesc *blocks;
    uint32_t guess, first_bad, ok_samples, c, i, used;
    size_t padded;
    LINNEApiResult result;

    /* block table in pinned host memory; grow until the hop fits */
    guess = block_limit ? block_limit : (uint32_t)((uint64_t)(data_size - start_offset) / 64u + 16u);
    if (!block_limit && h->num_samples_per_block) {
        const uint32_t by_samples = sample_limit / h->num_samples_per_block + 16u;
        if (by_samples < guess) guess = by_samples * 2u;
    }
    for (;;) {
        if (lnb_buf_reserve_host(&dec->h_blocks, (size_t)guess * sizeof(LnbBlockDesc))) return LINNE_APIRESULT_NG;
        blocks = (LnbBlockDesc *)dec->h_blocks.ptr;
        if (scan_blocks(h, data, data_size, start_offset, buffer_num_samples, sample_limit, block_limit,
                        blocks, guess, &scan) == 0) break;
        guess *= 2u;
    }
    if (scan_out) *scan_out = scan;
    if (first_bad_out) *first_bad_out = 0;

    if (consumed_bytes) *consumed_bytes = 0;
    if (decoded_samples) *decoded_samples = 0;
    if (scan.num_blocks == 0) return scan.framing_error;      /* OK when there was simply nothing to do */

    /* bytes of the image the blocks of this call span: a streaming caller passes everything that is left of its
     * stream with every DecodeBlock (tools/linne_player/linne_player.c:110-121) -- only this much is uploaded */
    used = scan.end_offset;
    for (i = 0; i < scan.num_blocks; i++) {
        const uint64_t end = (uint64_t)blocks[i].byte_off + blocks[i].byte_size;
        if (end > used) used = (end > data_size) ? data_size : (uint32_t)end;
    }
    if (!block_limit || d_stream_ext) used = data_size;

    /* device buffers */
    padded = LNB_ROUNDUP((size_t)used + 16u, 16u);
    lnb_fill_stream_cfg(&batch.cfg, h);
    {   /* the terminal block (post-CRC error) may not fit the PCM planes: park it inside the stride */
        uint32_t worst = scan.total_samples;
        if (scan.num_blocks > scan.num_decodable) worst += blocks[scan.num_blocks - 1].nsmp;
        batch.cfg.pcm_stride = (uint32_t)LNB_ROUNDUP((size_t)worst + 4u, 4u);
    }
    batch.cfg.work_stride = 0;
    batch.cfg.check_crc = (dec->flags & DEC_FLAG_CHECK_CRC) ? 1u : 0u;
    if (d_pcm_ext) batch.cfg.pcm_stride = pcm_stride_ext;
    if ((!d_stream_ext && lnb_buf_reserve_device(dec->dev, &dec->d_stream, padded))
        || lnb_buf_reserve_device(dec->dev, &dec->d_blocks, (size_t)scan.num_blocks * sizeof(LnbBlockDesc))
        || lnb_buf_reserve_device(dec->dev, &dec->d_params, (size_t)scan.num_blocks * C * sizeof(LnbChanParams))
        || (!d_pcm_ext && lnb_buf_reserve_device(dec->dev, &dec->d_pcm, (size_t)batch.cfg.pcm_stride * C * sizeof(int32_t)))
        || (staged && lnb_buf_reserve_host(&dec->h_pcm, (size_t)batch.cfg.pcm_stride * C * sizeof(int32_t))))
        return LINNE_APIRESULT_NG;

    /* a terminal block with a post-CRC error takes part in the CRC pass only */
    for (i = scan.num_decodable; i < scan.num_blocks; i++) blocks[i].type = 0xFFu;
    {   /* compressed blocks go to the fused streaming kernel; count what is left for the split kernels */
        const char *split = getenv("LINNE_B200_SPLIT_DECODE");
        batch.fused_max_n = (split && *split == '1') ? 0u : lnb_shim_fused_max_n();
        batch.num_plain_blocks = 0;
        batch.max_nsmp = 0;
        for (i = 0; i < scan.num_blocks; i++) {
            if (blocks[i].type != LNB_BLOCK_COMPRESSED || blocks[i].nsmp == 0u || blocks[i].nsmp > batch.fused_max_n)
                batch.num_plain_blocks++;
            if (blocks[i].nsmp > batch.max_nsmp) batch.max_nsmp = blocks[i].nsmp;
        }
    }

    {   /* Large batches: eight lanes per block / one per (block, channel) instead of one CTA per block (lnb_tput_v2.cuh) */
        const uint32_t min_blocks = dec->tput_min_blocks;
        const int32_t *pcm_base = d_pcm_ext ? d_pcm_ext : (const int32_t *)dec->d_pcm.ptr;
        batch.tput = (min_blocks && scan.num_blocks >= min_blocks && batch.fused_max_n
                      && ((uintptr_t)pcm_base & 15u) == 0u && (batch.cfg.pcm_stride & 3u) == 0u
                      && lnb_shim_tput_supported(&batch.cfg)) ? 1u : 0u;
    }

    batch.tab = *lnb_shim_tables(dec->dev);
    batch.stream = d_stream_ext ? d_stream_ext : (const uint8_t *)dec->d_stream.ptr;
    batch.stream_size = used;
    batch.blocks = (LnbBlockDesc *)dec->d_blocks.ptr;
    batch.num_blocks = scan.num_blocks;
    batch.params = (LnbChanParams *)dec->d_params.ptr;
    batch.pcm = d_pcm_ext ? d_pcm_ext : (int32_t *)dec->d_pcm.ptr;

    DEC_TRACE(dec, "hopped");
    if (!d_stream_ext) {
        if (dec->turn) lnb_turnstile_wait(dec->turn, 0, dec->turn_index);
        DEC_TRACE(dec, "upload turn");
        lnb_shim_memset(dec->dev, (uint8_t *)dec->d_stream.ptr + (padded - 16u), 0, 16u);
        lnb_shim_h2d(dec->dev, dec->d_stream.ptr, data, used);
    }
    lnb_shim_h2d(dec->dev, dec->d_blocks.ptr, blocks, (size_t)scan.num_blocks * sizeof(LnbBlockDesc));
    if (dec->turn && !d_stream_ext) {                            /* the next range of this device may upload while this one computes */
        const int bad = lnb_shim_sync(dec->dev);
        lnb_turnstile_pass(dec->turn, 0, dec->turn_index);
        DEC_TRACE(dec, "uploaded");
        if (bad) return LINNE_APIRESULT_NG;
    }
    if (lnb_shim_decode(dec->dev, &batch)) return LINNE_APIRESULT_NG;
    DEC_TRACE(dec, "launched");
    lnb_shim_d2h(dec->dev, blocks, dec->d_blocks.ptr, (size_t)scan.num_blocks * sizeof(LnbBlockDesc));
    if (staged && scan.total_samples) {
        /* pinned planes: the samples travel with the block table, the verdict below decides what is handed on */
        dec->stage_stride = batch.cfg.pcm_stride;
        for (c = 0; c < C; c++)
            lnb_shim_d2h(dec->dev, (int32_t *)dec->h_pcm.ptr + (size_t)c * batch.cfg.pcm_stride,
                         (int32_t *)dec->d_pcm.ptr + (size_t)c * batch.cfg.pcm_stride,
                         (size_t)scan.total_samples * sizeof(int32_t));
    }
    if (lnb_shim_sync(dec->dev)) return LINNE_APIRESULT_NG;
    DEC_TRACE(dec, "kernels done");

    /* first block (stream order) whose CRC failed outranks everything after it */
    first_bad = scan.num_blocks;
    if (dec->flags & DEC_FLAG_CHECK_CRC)
        for (i = 0; i < scan.num_blocks; i++) if (blocks[i].status & LNB_ST_CRC_MISMATCH) { first_bad = i; break; }

    if (first_bad < scan.num_blocks) {
        result = LINNE_APIRESULT_DETECT_DATA_CORRUPTION;
        ok_samples = blocks[first_bad].smp_off;
    } else if (scan.num_blocks > scan.num_decodable) {
        result = scan.post_crc_error;
        ok_samples = scan.total_samples;
        first_bad = scan.num_decodable;
    } else {
        result = scan.framing_error;
        ok_samples = scan.total_samples;
    }
    if (first_bad_out) *first_bad_out = first_bad;

    /* hand back every sample the reference would have produced before stopping */
    if (!d_pcm_ext && !staged) {
        int bad;
        if (dec->turn) lnb_turnstile_wait(dec->turn, 1, dec->turn_index);
        DEC_TRACE(dec, "download turn");
        for (c = 0; c < C; c++)
            lnb_shim_d2h(dec->dev, buffer[c], (int32_t *)dec->d_pcm.ptr + (size_t)c * batch.cfg.pcm_stride,
                         (size_t)ok_samples * sizeof(int32_t));
        bad = lnb_shim_sync(dec->dev);
        if (dec->turn) lnb_turnstile_pass(dec->turn, 1, dec->turn_index);
        DEC_TRACE(dec, "downloaded");
        if (bad) return LINNE_APIRESULT_NG;
    }

    if (consumed_bytes) *consumed_bytes = scan.end_offset - start_offset;
    if (decoded_samples) *decoded_samples = ok_samples;
    return result;
}

/* ---- DecodeBlock: batch of one, or read-ahead ---------------------------------------------
 * A streaming caller (tools/linne_player/linne_player.c:110-121) asks for one block per call and passes all
 * that is left of its stream.  With read-ahead K the first call decodes the next K blocks in ONE batch
 * (K CTAs instead of one: the kernel is latency-bound per block, so K blocks take about as long as one) and
 * keeps their PCM in pinned host memory; the following calls are served from there.  A hit requires the
 * caller's bytes to equal the bytes that were decoded (memcmp of the whole block), so the result is the
 * function of (header, block bytes) it is in the reference, whatever the caller did in between; only
 * blocks that decoded cleanly are kept, every error takes the ordinary batch-of-one path. */
static int serve_cached_block(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
                              int32_t **buffer, uint32_t buffer_num_samples, uint32_t *decode_size, uint32_t *num_decode_samples)
{
    const struct LnbCachedBlock *b;
    uint32_t c;
    if (dec->ra_next >= dec->ra_count) return 0;
    b = &dec->ra_blocks[dec->ra_next];
    if (data_size < b->byte_size || b->nsmp > buffer_num_samples
        || memcmp(data, dec->ra_image + b->byte_off, b->byte_size) != 0) {
        dec->ra_count = dec->ra_next = 0;
        return 0;
    }
    for (c = 0; c < dec->header.num_channels; c++)
        memcpy(buffer[c], (const int32_t *)dec->h_pcm.ptr + (size_t)c * dec->stage_stride + b->smp_off, (size_t)b->nsmp * sizeof(int32_t));
    *decode_size = b->byte_size;
    *num_decode_samples = b->nsmp;
    dec->ra_next++;
    return 1;
}

static void fill_block_cache(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size)
{
    const LnbBlockDesc *blocks;
    BlockScan scan;
    uint32_t first_bad = 0, i, n, span;
    dec->ra_count = dec->ra_next = 0;
    memset(&scan, 0, sizeof(scan));
    if (decode_range(dec, data, data_size, 0, NULL, 0xFFFFFFFFu, LNB_READAHEAD_MAX_SAMPLES, dec->readahead, 1,
                     NULL, NULL, NULL, NULL, 0, &scan, &first_bad) == LINNE_APIRESULT_NG) return;
    blocks = (const LnbBlockDesc *)dec->h_blocks.ptr;
    if (blocks == NULL) return;
    n = (first_bad < scan.num_decodable) ? first_bad : scan.num_decodable;
    /* a block whose payload does not end where its size field says is left to the ordinary path */
    for (i = 0; i < n; i++) {
        uint32_t walked = LNB_BLOCK_HEADER_SIZE;
        if (blocks[i].type == LNB_BLOCK_COMPRESSED) walked += blocks[i].na;
        else if (blocks[i].type == LNB_BLOCK_RAW) walked += (dec->header.bits_per_sample / 8u) * blocks[i].nsmp * dec->header.num_channels;
        if (walked != blocks[i].byte_size) { n = i; break; }
    }
    if (n < 2) return;                                         /* nothing gained over a batch of one */
    span = blocks[n - 1].byte_off + blocks[n - 1].byte_size;
    if (span > dec->ra_image_cap) {
        uint8_t *p = (uint8_t *)realloc(dec->ra_image, span);
        if (!p) return;
        dec->ra_image = p; dec->ra_image_cap = span;
    }
    if (n > dec->ra_blocks_cap) {
        struct LnbCachedBlock *p = (struct LnbCachedBlock *)realloc(dec->ra_blocks, (size_t)n * sizeof(*p));
        if (!p) return;
        dec->ra_blocks = p; dec->ra_blocks_cap = n;
    }
    memcpy(dec->ra_image, data, span);
    for (i = 0; i < n; i++) {
        dec->ra_blocks[i].byte_off = blocks[i].byte_off; dec->ra_blocks[i].byte_size = blocks[i].byte_size;
        dec->ra_blocks[i].smp_off = blocks[i].smp_off; dec->ra_blocks[i].nsmp = blocks[i].nsmp;
    }
    dec->ra_count = n;
}

/* reference linne_decoder.c:564-668 */
LINNEApiResult LINNEDecoder_DecodeBlock(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
        int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples,
        uint32_t *decode_size, uint32_t *num_decode_samples)
{
    LINNEApiResult ret;
    uint32_t c, ok_samples = 0, consumed = 0;
    if (dec == NULL || data == NULL || buffer == NULL || decode_size == NULL || num_decode_samples == NULL)
        return LINNE_APIRESULT_INVALID_ARGUMENT;
    if (!(dec->flags & DEC_FLAG_HEADER_SET)) return LINNE_APIRESULT_PARAMETER_NOT_SET;
    if (buffer_num_channels < dec->header.num_channels) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    for (c = 0; c < dec->header.num_channels; c++) if (buffer[c] == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if (data_size == 0) return LINNE_APIRESULT_INSUFFICIENT_DATA;

    if (dec->readahead > 1u) {
        if (serve_cached_block(dec, data, data_size, buffer, buffer_num_samples, decode_size, num_decode_samples))
            return LINNE_APIRESULT_OK;
        fill_block_cache(dec, data, data_size);
        if (serve_cached_block(dec, data, data_size, buffer, buffer_num_samples, decode_size, num_decode_samples))
            return LINNE_APIRESULT_OK;
        dec->ra_count = dec->ra_next = 0;
    }
    ret = decode_range(dec, data, data_size, 0, NULL, buffer_num_samples, 0xFFFFFFFFu, 1, 1,
                       &consumed, &ok_samples, NULL, NULL, 0, NULL, NULL);
    for (c = 0; ok_samples && c < dec->header.num_channels; c++)
        memcpy(buffer[c], (const int32_t *)dec->h_pcm.ptr + (size_t)c * dec->stage_stride, (size_t)ok_samples * sizeof(int32_t));
    if (ret == LINNE_APIRESULT_OK) {
        /* the bytes the payload reader walked over (linne_decoder.c:655-661) */
        const LnbBlockDesc *blk = (const LnbBlockDesc *)dec->h_blocks.ptr;
        consumed = LNB_BLOCK_HEADER_SIZE + blk[0].na;
    }
    *decode_size = consumed;
    *num_decode_samples = ok_samples;
    return ret;
}


/* ---- several GPUs behind one handle -----------------------------------------------------------------------------
 * The hop over the size fields gives the block table; contiguous block ranges go to one child handle each (own device,
 * own host thread): a range is a stream of its own for the child -- it uploads the range's bytes, decodes, and copies
 * its samples to their place in the caller's planes.  Outputs are disjoint, nothing is exchanged.  The call returns
 * the first error in stream order; samples behind a failing block may have been written by later ranges (the reference
 * leaves them untouched). */
void LINNEB200_DecoderSetDevices(struct LINNEDecoder *dec, uint32_t num_devices)
{
    if (dec == NULL) return;
    if (num_devices > LNB_MAX_DEVICES) num_devices = LNB_MAX_DEVICES;
    dec->num_devices = num_devices;
}

struct LnbDecShard {
    struct LINNEDecoder *child;
    const uint8_t *data; uint32_t data_size;
    int32_t *planes[LINNE_MAX_NUM_CHANNELS];   /* int32 planes, or ... */
    uint8_t *packed;                       /* ... interleaved packed PCM for this range */
    uint32_t room, sample_limit, block_limit;
    uint32_t first_sample;                 /* of the range, in the whole stream */
    uint32_t decoded;
    LINNEApiResult result;
    int ordinal;
};

static LINNEApiResult decode_packed_range(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size, uint32_t start_offset,
        uint8_t *pcm, uint32_t room_frames, uint32_t sample_limit, uint32_t block_limit, uint32_t *decoded_out);

static void *dec_shard_main(void *arg)
{
    struct LnbDecShard *sh = (struct LnbDecShard *)arg;
    lnb_shim_set_device(sh->ordinal);
    sh->decoded = 0;
    if (sh->packed) sh->result = decode_packed_range(sh->child, sh->data, sh->data_size, 0, sh->packed, sh->room, sh->sample_limit, sh->block_limit, &sh->decoded);
    else sh->result = decode_range(sh->child, sh->data, sh->data_size, 0, sh->planes, sh->room, sh->sample_limit, sh->block_limit, 0,
                                   NULL, &sh->decoded, NULL, NULL, 0, NULL, NULL);
    if (sh->child->turn) {                                       /* whatever happened: nobody waits for this range any more */
        lnb_turnstile_pass(sh->child->turn, 0, sh->child->turn_index);
        lnb_turnstile_pass(sh->child->turn, 1, sh->child->turn_index);
        sh->child->turn = NULL;
    }
    return NULL;
}

/* DecodeWhole over dec->num_devices block ranges; returns 0 when the stream is too short to be worth it (*ret untouched) */
static int decode_whole_sharded(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
                                int32_t **buffer, uint8_t *packed, uint32_t buffer_num_samples, LINNEApiResult *ret, uint32_t *decoded_total)
{
    const struct LINNEHeader *h = &dec->header;
    const int ndev = lnb_shim_device_count();
    struct LnbDecShard sh[LNB_MAX_DEVICES];
    pthread_t th[LNB_MAX_DEVICES];
    LnbTurnstile turn;
    BlockScan scan;
    uint32_t guess, G, k, c, started = 0, ndev_used = 1;
    int home, own;
    if (ndev <= 0) return 0;
    guess = (uint32_t)((uint64_t)(data_size - LINNE_HEADER_SIZE) / 64u + 16u);
    if (h->num_samples_per_block) {
        const uint32_t by_samples = h->num_samples / h->num_samples_per_block + 16u;
        if (by_samples < guess) guess = by_samples * 2u;
    }
    for (;;) {
        if (guess > dec->shard_table_cap) {
            LnbBlockDesc *p = (LnbBlockDesc *)realloc(dec->shard_table, (size_t)guess * sizeof(LnbBlockDesc));
            if (!p) return 0;
            dec->shard_table = p; dec->shard_table_cap = guess;
        }
        if (scan_blocks(h, data, data_size, LINNE_HEADER_SIZE, buffer_num_samples, h->num_samples, 0,
                        dec->shard_table, dec->shard_table_cap, &scan) == 0) break;
        guess *= 2u;
    }
    /* ranges per requested device (pipelining), devices shared round-robin when fewer are visible than asked for */
    G = lnb_plan_ranges(scan.num_decodable, dec->num_devices > 1u ? dec->num_devices : 1u);
    ndev_used = dec->num_devices > 1u ? (dec->num_devices < (uint32_t)ndev ? dec->num_devices : (uint32_t)ndev) : 1u;
    if (G < 2u) return 0;                                        /* a handful of blocks: the handle's own device */
    home = lnb_shim_current_device();
    own = lnb_shim_device_ordinal(dec->dev);
    for (k = 0; k < G; k++) {
        if (!dec->child[k]) {
            lnb_shim_set_device(dec->num_devices > 1u ? (int)(k % ndev_used) : own);
            dec->child[k] = LINNEDecoder_Create(&dec->config, NULL, 0);
            if (dec->child[k]) dec->child[k]->num_devices = 0;
        }
        if (dec->child[k]) {
            /* the kernels of an earlier range of a device go first wherever SMs free up: its download can start while the
             * later ranges still compute (without this all ranges' kernels end together and the link idles until then) */
            const uint32_t depth = (G + ndev_used - 1u) / ndev_used, place = k / ndev_used;
            dec->child[k]->rank_override = depth > 1u ? (int)(7u - (place * 7u) / (depth - 1u)) : -1;
        }
        if (!dec->child[k] || LINNEDecoder_SetHeader(dec->child[k], h) != LINNE_APIRESULT_OK) { if (home >= 0) lnb_shim_set_device(home); return 0; }
        /* ranges of one call keep each other's kernels company: the throughput kernels pay from ~1000 blocks on then
         * (include/linne_b200.h: LINNEB200_DecoderSetThroughputBlocks) */
        dec->child[k]->tput_min_blocks = (dec->tput_min_blocks > 1024u) ? 1024u : dec->tput_min_blocks;
    }
    if (home >= 0) lnb_shim_set_device(home);
    lnb_turnstile_init(&turn, dec->num_devices > 1u ? ndev_used : 1u);
    for (k = 0; k < G; k++) {
        /* contiguous ranges of the decodable blocks; the last range runs to the end of the data and meets whatever ends
         * the stream (framing error, terminal block) exactly as the single-device path does */
        const uint32_t b0 = (uint32_t)((uint64_t)scan.num_decodable * k / G), b1 = (uint32_t)((uint64_t)scan.num_decodable * (k + 1u) / G);
        const LnbBlockDesc *first = &dec->shard_table[b0];
        const int last = (k + 1u == G);
        const uint32_t end_byte = last ? data_size : dec->shard_table[b1].byte_off;
        memset(&sh[k], 0, sizeof(sh[k]));
        sh[k].child = dec->child[k];
        sh[k].data = data + first->byte_off;
        sh[k].data_size = end_byte - first->byte_off;
        for (c = 0; c < h->num_channels && buffer; c++) sh[k].planes[c] = buffer[c] + first->smp_off;
        if (packed) sh[k].packed = packed + (size_t)first->smp_off * h->num_channels * (h->bits_per_sample / 8u);
        sh[k].first_sample = first->smp_off;
        sh[k].room = buffer_num_samples - first->smp_off;
        sh[k].sample_limit = last ? h->num_samples - first->smp_off : dec->shard_table[b1].smp_off - first->smp_off;
        sh[k].block_limit = last ? 0u : b1 - b0;
        sh[k].ordinal = dec->num_devices > 1u ? (int)(k % ndev_used) : own;
        dec->child[k]->turn = &turn; dec->child[k]->turn_index = k;
        if (pthread_create(&th[k], NULL, dec_shard_main, &sh[k]) != 0) {
            dec->child[k]->turn = NULL;
            lnb_turnstile_pass(&turn, 0, k); lnb_turnstile_pass(&turn, 1, k);
            break;
        }
        started++;
    }
    for (k = 0; k < started; k++) pthread_join(th[k], NULL);
    lnb_turnstile_destroy(&turn);
    if (started < G) { *ret = LINNE_APIRESULT_NG; return 1; }
    *ret = LINNE_APIRESULT_OK;
    if (decoded_total) *decoded_total = 0;
    for (k = 0; k < G; k++) {                                    /* first error in stream order; frames up to there count */
        if (decoded_total) *decoded_total = sh[k].first_sample + sh[k].decoded;
        if (sh[k].result != LINNE_APIRESULT_OK) { *ret = sh[k].result; break; }
    }
    return 1;
}

/* reference linne_decoder.c:671-730 */
LINNEApiResult LINNEDecoder_DecodeWhole(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
        int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples)
{
    struct LINNEHeader header;
    LINNEApiResult ret;
    uint32_t c;
    if (dec == NULL || data == NULL || buffer == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if ((ret = LINNEDecoder_DecodeHeader(data, data_size, &header)) != LINNE_APIRESULT_OK) return ret;
    if ((ret = LINNEDecoder_SetHeader(dec, &header)) != LINNE_APIRESULT_OK) return ret;
    if (buffer_num_channels < header.num_channels || buffer_num_samples < header.num_samples)
        return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    for (c = 0; c < header.num_channels; c++) if (buffer[c] == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    /* several devices, or a stream long enough to pipeline its transfers against its kernels on one */
    if ((dec->num_devices > 1u || data_size >= (32u << 20)) && decode_whole_sharded(dec, data, data_size, buffer, NULL, buffer_num_samples, &ret, NULL)) return ret;
    return decode_range(dec, data, data_size, LINNE_HEADER_SIZE, buffer, buffer_num_samples,
                        header.num_samples, 0, 0, NULL, NULL, NULL, NULL, 0, NULL, NULL);
}

LnbDevice *lnb_decoder_device(const struct LINNEDecoder *dec) { return dec->dev; }

void lnb_decoder_set_tput_min_blocks(struct LINNEDecoder *dec, uint32_t blocks) { dec->tput_min_blocks = blocks; }

void lnb_decoder_set_readahead(struct LINNEDecoder *dec, uint32_t blocks)
{
    dec->readahead = (blocks > LNB_MAX_READAHEAD) ? LNB_MAX_READAHEAD : blocks;
    dec->ra_count = dec->ra_next = 0;
}

/* ---- extension entry point (include/linne_b200.h) ---- */
#include "linne_b200.h"

static LINNEApiResult decode_files(struct LINNEDecoder *dec, const uint8_t *host_image, const uint8_t *d_data, uint32_t data_size,
        struct LINNEB200FileDesc *files, uint32_t num_files, int32_t *d_pcm, uint32_t pcm_stride, uint32_t max_channels);

LINNEApiResult LINNEB200_DecodeWholeResident(struct LINNEDecoder *dec, const uint8_t *data, const uint8_t *d_data,
        uint32_t data_size, int32_t *d_pcm, uint32_t pcm_stride, uint32_t buffer_num_channels, uint32_t buffer_num_samples)
{
    struct LINNEHeader header;
    LINNEApiResult ret;
    if (dec == NULL || d_data == NULL || d_pcm == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if (data == NULL) {
        /* no host copy supplied: the hop over the size fields runs on the device (one stream of a corpus batch) */
        struct LINNEB200FileDesc one;
        if (pcm_stride < buffer_num_samples) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
        one.first_sample = 0; one.num_samples = buffer_num_samples; one.out_offset = 0; one.out_size = data_size; one.status = 0;
        (void)header;
        return decode_files(dec, NULL, d_data, data_size, &one, 1, d_pcm, pcm_stride, buffer_num_channels);
    }
    if ((ret = LINNEDecoder_DecodeHeader(data, data_size, &header)) != LINNE_APIRESULT_OK) return ret;
    if ((ret = LINNEDecoder_SetHeader(dec, &header)) != LINNE_APIRESULT_OK) return ret;
    if (buffer_num_channels < header.num_channels || buffer_num_samples < header.num_samples
        || pcm_stride < buffer_num_samples) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    return decode_range(dec, data, data_size, LINNE_HEADER_SIZE, NULL, buffer_num_samples,
                        header.num_samples, 0, 0, NULL, NULL, d_data, d_pcm, pcm_stride, NULL, NULL);
}

/* Several streams per call (corpus batches): the streams sit one after the other in the device image
 * `d_data` (files[i].out_offset / out_size), all with the same stream parameters; the blocks of all of them go
 * through the kernels as ONE batch and every file's PCM lands at files[i].first_sample of the planes. */
static LINNEApiResult decode_files(struct LINNEDecoder *dec, const uint8_t *host_image, const uint8_t *d_data, uint32_t data_size,
        struct LINNEB200FileDesc *files, uint32_t num_files, int32_t *d_pcm, uint32_t pcm_stride, uint32_t max_channels)
{
    struct LINNEHeader h0, h1;
    LnbDecodeBatch batch;
    LnbBlockDesc *blocks;
    const uint8_t *img;
    LINNEApiResult ret, overall = LINNE_APIRESULT_OK;
    uint32_t i, k, nb = 0, cap, C = 0;
    uint32_t *first_block;
    int hopped = 0;
    if (dec == NULL || d_data == NULL || files == NULL || num_files == 0 || d_pcm == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    for (i = 0; i < num_files; i++) {
        if ((uint64_t)files[i].out_offset + files[i].out_size > data_size
            || (uint64_t)files[i].first_sample + files[i].num_samples > pcm_stride) return LINNE_APIRESULT_INVALID_ARGUMENT;
        files[i].status = (int32_t)LINNE_APIRESULT_OK;
    }
    img = host_image;
    if (!(first_block = (uint32_t *)malloc(((size_t)num_files + 1u) * sizeof(uint32_t)))) return LINNE_APIRESULT_NG;
    if (!host_image) {
        /* The image is in HBM only: the hop over the size fields runs there (lnb_hop.cuh).  What comes back is a result
         * record per stream (with its 30 header bytes) and the block table -- not the image. */
        LnbHopFile *hf;
        LnbHopResult *hr;
        LnbBlockDesc *ht;
        uint32_t total_cap = 0;
        size_t staging;
        int fallback = 0;
        /* table slices: a conforming stream has one block per num_samples_per_block frames; the frames the caller gave
         * room for bound that from above for any block size >= 256 (smaller: the hop reports the overflow) */
        for (i = 0; i < num_files; i++) total_cap += files[i].num_samples / 256u + 64u;
        staging = (size_t)num_files * (sizeof(LnbHopFile) + sizeof(LnbHopResult)) + (size_t)total_cap * sizeof(LnbBlockDesc);
        if (lnb_buf_reserve_host(&dec->h_hop, staging)
            || lnb_buf_reserve_device(dec->dev, &dec->d_hop_files, (size_t)num_files * sizeof(LnbHopFile))
            || lnb_buf_reserve_device(dec->dev, &dec->d_hop_res, (size_t)num_files * sizeof(LnbHopResult))
            || lnb_buf_reserve_device(dec->dev, &dec->d_hop_table, (size_t)total_cap * sizeof(LnbBlockDesc))
            || lnb_buf_reserve_host(&dec->h_blocks, (size_t)total_cap * sizeof(LnbBlockDesc))) { free(first_block); return LINNE_APIRESULT_NG; }
        hf = (LnbHopFile *)dec->h_hop.ptr;
        hr = (LnbHopResult *)(hf + num_files);
        ht = (LnbBlockDesc *)(hr + num_files);
        total_cap = 0;
        for (i = 0; i < num_files; i++) {
            hf[i].offset = files[i].out_offset; hf[i].size = files[i].out_size; hf[i].room_samples = files[i].num_samples;
            hf[i].table_first = total_cap; hf[i].table_cap = files[i].num_samples / 256u + 64u;
            total_cap += hf[i].table_cap;
        }
        lnb_shim_h2d(dec->dev, dec->d_hop_files.ptr, hf, (size_t)num_files * sizeof(LnbHopFile));
        if (lnb_shim_hop(dec->dev, d_data, (const LnbHopFile *)dec->d_hop_files.ptr, num_files,
                         (LnbBlockDesc *)dec->d_hop_table.ptr, (LnbHopResult *)dec->d_hop_res.ptr)) { free(first_block); return LINNE_APIRESULT_NG; }
        lnb_shim_d2h(dec->dev, hr, dec->d_hop_res.ptr, (size_t)num_files * sizeof(LnbHopResult));
        if (lnb_shim_sync(dec->dev)) { free(first_block); return LINNE_APIRESULT_NG; }
        for (i = 0; i < num_files; i++) if (hr[i].overflow) fallback = 1;
        if (!fallback) {
            /* only the slices that were filled travel */
            for (i = 0; i < num_files; i++)
                if (hr[i].num_blocks) lnb_shim_d2h(dec->dev, ht + hf[i].table_first, (LnbBlockDesc *)dec->d_hop_table.ptr + hf[i].table_first,
                                                   (size_t)hr[i].num_blocks * sizeof(LnbBlockDesc));
            if (lnb_shim_sync(dec->dev)) { free(first_block); return LINNE_APIRESULT_NG; }
            if ((ret = LINNEDecoder_DecodeHeader(hr[0].header, files[0].out_size < 32u ? files[0].out_size : 32u, &h0)) != LINNE_APIRESULT_OK
                || (ret = LINNEDecoder_SetHeader(dec, &h0)) != LINNE_APIRESULT_OK) { free(first_block); return ret; }
            C = h0.num_channels;
            if (max_channels && max_channels < C) { free(first_block); return LINNE_APIRESULT_INSUFFICIENT_BUFFER; }
            blocks = (LnbBlockDesc *)dec->h_blocks.ptr;
            for (i = 0; i < num_files; i++) {
                first_block[i] = nb;
                ret = LINNEDecoder_DecodeHeader(hr[i].header, files[i].out_size < 32u ? files[i].out_size : 32u, &h1);
                if (ret == LINNE_APIRESULT_OK
                    && (h1.num_channels != h0.num_channels || h1.bits_per_sample != h0.bits_per_sample || h1.preset != h0.preset
                        || h1.num_samples_per_block != h0.num_samples_per_block || h1.ch_process_method != h0.ch_process_method
                        || h1.format_version != h0.format_version || h1.codec_version != h0.codec_version))
                    ret = LINNE_APIRESULT_INVALID_FORMAT;                     /* one set of stream parameters per batch */
                if (ret == LINNE_APIRESULT_OK && files[i].num_samples < h1.num_samples) ret = LINNE_APIRESULT_INSUFFICIENT_BUFFER;
                if (ret != LINNE_APIRESULT_OK) { files[i].status = (int32_t)ret; continue; }
                memcpy(blocks + nb, ht + hf[i].table_first, (size_t)hr[i].num_blocks * sizeof(LnbBlockDesc));
                for (k = 0; k < hr[i].num_blocks; k++) blocks[nb + k].smp_off += files[i].first_sample;
                for (k = hr[i].num_decodable; k < hr[i].num_blocks; k++) blocks[nb + k].type = 0xFFu;    /* CRC pass only */
                if (hr[i].num_blocks > hr[i].num_decodable) files[i].status = (int32_t)hr[i].post_crc_error;
                else files[i].status = (int32_t)hr[i].framing_error;
                nb += hr[i].num_blocks;
            }
            hopped = 1;
        } else {
            /* a stream of tiny blocks: the image comes to the host once and the hop runs there */
            if (lnb_buf_reserve_host(&dec->h_stream, (size_t)data_size + 16u)) { free(first_block); return LINNE_APIRESULT_NG; }
            lnb_shim_d2h(dec->dev, dec->h_stream.ptr, d_data, data_size);
            if (lnb_shim_sync(dec->dev)) { free(first_block); return LINNE_APIRESULT_NG; }
            img = (const uint8_t *)dec->h_stream.ptr;
        }
    }
    if (!hopped) {
    if ((ret = LINNEDecoder_DecodeHeader(img + files[0].out_offset, files[0].out_size, &h0)) != LINNE_APIRESULT_OK) { free(first_block); return ret; }
    if ((ret = LINNEDecoder_SetHeader(dec, &h0)) != LINNE_APIRESULT_OK) { free(first_block); return ret; }
    C = h0.num_channels;
    if (max_channels && max_channels < C) { free(first_block); return LINNE_APIRESULT_INSUFFICIENT_BUFFER; }

    cap = 64u;
    for (i = 0; i < num_files; i++) cap += files[i].out_size / 64u + 16u;
    if (lnb_buf_reserve_host(&dec->h_blocks, (size_t)cap * sizeof(LnbBlockDesc))) { free(first_block); return LINNE_APIRESULT_NG; }
    blocks = (LnbBlockDesc *)dec->h_blocks.ptr;
    for (i = 0; i < num_files; i++) {
        BlockScan scan;
        first_block[i] = nb;
        ret = LINNEDecoder_DecodeHeader(img + files[i].out_offset, files[i].out_size, &h1);
        if (ret == LINNE_APIRESULT_OK
            && (h1.num_channels != h0.num_channels || h1.bits_per_sample != h0.bits_per_sample || h1.preset != h0.preset
                || h1.num_samples_per_block != h0.num_samples_per_block || h1.ch_process_method != h0.ch_process_method
                || h1.format_version != h0.format_version || h1.codec_version != h0.codec_version))
            ret = LINNE_APIRESULT_INVALID_FORMAT;                     /* one set of stream parameters per batch */
        if (ret == LINNE_APIRESULT_OK && files[i].num_samples < h1.num_samples) ret = LINNE_APIRESULT_INSUFFICIENT_BUFFER;
        if (ret != LINNE_APIRESULT_OK) { files[i].status = (int32_t)ret; continue; }
        /* offsets relative to the image; a file's hop cannot run into its neighbour */
        if (scan_blocks(&h1, img, files[i].out_offset + files[i].out_size, files[i].out_offset + LINNE_HEADER_SIZE,
                        files[i].num_samples, h1.num_samples, 0, blocks + nb, cap - nb, &scan)) {
            free(first_block);
            return LINNE_APIRESULT_NG;                                /* the table bound above covers every legal stream */
        }
        for (k = 0; k < scan.num_blocks; k++) blocks[nb + k].smp_off += files[i].first_sample;
        for (k = scan.num_decodable; k < scan.num_blocks; k++) blocks[nb + k].type = 0xFFu;    /* CRC pass only */
        if (scan.num_blocks > scan.num_decodable) files[i].status = (int32_t)scan.post_crc_error;
        else files[i].status = (int32_t)scan.framing_error;
        nb += scan.num_blocks;
    }
    }
    first_block[num_files] = nb;

    if (nb) {
        lnb_fill_stream_cfg(&batch.cfg, &h0);
        batch.cfg.pcm_stride = pcm_stride;
        batch.cfg.work_stride = 0;
        batch.cfg.check_crc = (dec->flags & DEC_FLAG_CHECK_CRC) ? 1u : 0u;
        if (lnb_buf_reserve_device(dec->dev, &dec->d_blocks, (size_t)nb * sizeof(LnbBlockDesc))
            || lnb_buf_reserve_device(dec->dev, &dec->d_params, (size_t)nb * C * sizeof(LnbChanParams))) { free(first_block); return LINNE_APIRESULT_NG; }
        {
            const char *split = getenv("LINNE_B200_SPLIT_DECODE");
            batch.fused_max_n = (split && *split == '1') ? 0u : lnb_shim_fused_max_n();
            batch.num_plain_blocks = 0;
            batch.max_nsmp = 0;
            for (k = 0; k < nb; k++) {
                if (blocks[k].type != LNB_BLOCK_COMPRESSED || blocks[k].nsmp == 0u || blocks[k].nsmp > batch.fused_max_n)
                    batch.num_plain_blocks++;
                if (blocks[k].nsmp > batch.max_nsmp) batch.max_nsmp = blocks[k].nsmp;
            }
        }
        batch.tput = (dec->tput_min_blocks && nb >= dec->tput_min_blocks && batch.fused_max_n
                      && ((uintptr_t)d_pcm & 15u) == 0u && (pcm_stride & 3u) == 0u
                      && lnb_shim_tput_supported(&batch.cfg)) ? 1u : 0u;
        batch.tab = *lnb_shim_tables(dec->dev);
        batch.stream = d_data;
        batch.stream_size = data_size;
        batch.blocks = (LnbBlockDesc *)dec->d_blocks.ptr;
        batch.num_blocks = nb;
        batch.params = (LnbChanParams *)dec->d_params.ptr;
        batch.pcm = d_pcm;
        lnb_shim_h2d(dec->dev, dec->d_blocks.ptr, blocks, (size_t)nb * sizeof(LnbBlockDesc));
        if (lnb_shim_decode(dec->dev, &batch)) { free(first_block); return LINNE_APIRESULT_NG; }
        lnb_shim_d2h(dec->dev, blocks, dec->d_blocks.ptr, (size_t)nb * sizeof(LnbBlockDesc));
        if (lnb_shim_sync(dec->dev)) { free(first_block); return LINNE_APIRESULT_NG; }
    }
    for (i = 0; i < num_files; i++) {
        if (dec->flags & DEC_FLAG_CHECK_CRC)
            for (k = first_block[i]; k < first_block[i + 1u]; k++)
                if (blocks[k].status & LNB_ST_CRC_MISMATCH) { files[i].status = (int32_t)LINNE_APIRESULT_DETECT_DATA_CORRUPTION; break; }
        if (overall == LINNE_APIRESULT_OK && files[i].status != (int32_t)LINNE_APIRESULT_OK) overall = (LINNEApiResult)files[i].status;
    }
    free(first_block);
    return overall;
}

LINNEApiResult LINNEB200_DecodeFilesResident(struct LINNEDecoder *dec, const uint8_t *d_data, uint32_t data_size,
        struct LINNEB200FileDesc *files, uint32_t num_files, int32_t *d_pcm, uint32_t pcm_stride)
{
    return decode_files(dec, NULL, d_data, data_size, files, num_files, d_pcm, pcm_stride, 0u);
}

/* The same for host buffers: the streams in the host image `data` (files[i].out_offset / out_size), the PCM of the
 * files back to back as packed interleaved samples in `pcm`, files[i].num_samples frames of room each
 * (first_sample is filled in: the frame offset of the file in `pcm`).  One transfer up, one batch, one down. */
LINNEApiResult LINNEB200_DecodeFilesPacked(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
        struct LINNEB200FileDesc *files, uint32_t num_files, uint8_t *pcm)
{
    struct LINNEHeader h0;
    LINNEApiResult ret;
    uint64_t total = 0;
    uint32_t i, bytes, C;
    size_t stride, padded;
    if (dec == NULL || data == NULL || files == NULL || num_files == 0 || pcm == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if (files[0].out_offset > data_size || files[0].out_size > data_size - files[0].out_offset) return LINNE_APIRESULT_INVALID_ARGUMENT;
    if ((ret = LINNEDecoder_DecodeHeader(data + files[0].out_offset, files[0].out_size, &h0)) != LINNE_APIRESULT_OK) return ret;
    bytes = h0.bits_per_sample / 8u; C = h0.num_channels;
    if (bytes == 0 || bytes > 4u || (h0.bits_per_sample % 8u) != 0u || C == 0u) return LINNE_APIRESULT_INVALID_FORMAT;
    for (i = 0; i < num_files; i++) { files[i].first_sample = (uint32_t)total; total += files[i].num_samples; }
    if (total == 0 || total > 0xFFFFFFF0ull) return LINNE_APIRESULT_INVALID_ARGUMENT;
    stride = LNB_ROUNDUP((size_t)total + 4u, 4u);
    padded = LNB_ROUNDUP((size_t)data_size + 16u, 16u);
    if (lnb_buf_reserve_device(dec->dev, &dec->d_stream, padded)
        || lnb_buf_reserve_device(dec->dev, &dec->d_pcm, stride * C * sizeof(int32_t))
        || lnb_buf_reserve_device(dec->dev, &dec->d_packed, (size_t)total * C * bytes + 16u)) return LINNE_APIRESULT_NG;
    lnb_shim_memset(dec->dev, (uint8_t *)dec->d_stream.ptr + (padded - 16u), 0, 16u);
    lnb_shim_h2d(dec->dev, dec->d_stream.ptr, data, data_size);
    ret = decode_files(dec, data, (const uint8_t *)dec->d_stream.ptr, data_size, files, num_files, (int32_t *)dec->d_pcm.ptr, (uint32_t)stride, 0u);
    /* every file that decoded is handed back; a failed file's frames are undefined */
    if (lnb_shim_pack_pcm(dec->dev, (const int32_t *)dec->d_pcm.ptr, (uint8_t *)dec->d_packed.ptr, (uint32_t)stride,
                          (uint32_t)total, C, bytes)) return LINNE_APIRESULT_NG;
    lnb_shim_d2h(dec->dev, pcm, dec->d_packed.ptr, (size_t)total * C * bytes);
    if (lnb_shim_sync(dec->dev)) return LINNE_APIRESULT_NG;
    return ret;
}

/* Packed interleaved PCM out (the bytes of a WAV data chunk), converted from the planes on the device: the blocks found
 * from `start_offset` on (header already set), at most block_limit blocks (0: all) / sample_limit frames. */
static LINNEApiResult decode_packed_range(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size, uint32_t start_offset,
        uint8_t *pcm, uint32_t room_frames, uint32_t sample_limit, uint32_t block_limit, uint32_t *decoded_out)
{
    const struct LINNEHeader *h = &dec->header;
    const uint32_t bytes = h->bits_per_sample / 8u;
    const uint32_t frames = sample_limit < room_frames ? sample_limit : room_frames;
    LINNEApiResult ret;
    uint32_t decoded = 0;
    size_t stride = LNB_ROUNDUP((size_t)frames + (size_t)h->num_samples_per_block + 65536u + 4u, 4u);   /* room for a parked terminal block */
    if (lnb_buf_reserve_device(dec->dev, &dec->d_pcm, stride * h->num_channels * sizeof(int32_t))
        || lnb_buf_reserve_device(dec->dev, &dec->d_packed, (size_t)frames * h->num_channels * bytes + 16u))
        return LINNE_APIRESULT_NG;
    ret = decode_range(dec, data, data_size, start_offset, NULL, room_frames, sample_limit, block_limit, 0,
                       NULL, &decoded, NULL, (int32_t *)dec->d_pcm.ptr, (uint32_t)stride, NULL, NULL);
    if (decoded) {          /* every sample the reference would have produced before stopping */
        if (lnb_shim_pack_pcm(dec->dev, (const int32_t *)dec->d_pcm.ptr, (uint8_t *)dec->d_packed.ptr, (uint32_t)stride,
                              decoded, h->num_channels, bytes)) return LINNE_APIRESULT_NG;
        if (dec->turn) {                                         /* kernels done before this range takes its turn on the link */
            if (lnb_shim_sync(dec->dev)) return LINNE_APIRESULT_NG;
            lnb_turnstile_wait(dec->turn, 1, dec->turn_index);
        }
        lnb_shim_d2h(dec->dev, pcm, dec->d_packed.ptr, (size_t)decoded * h->num_channels * bytes);
        {
            const int bad = lnb_shim_sync(dec->dev);
            if (dec->turn) lnb_turnstile_pass(dec->turn, 1, dec->turn_index);
            if (bad) return LINNE_APIRESULT_NG;
        }
    }
    *decoded_out = decoded;
    return ret;
}

LINNEApiResult LINNEB200_DecodeWholePacked(struct LINNEDecoder *dec, const uint8_t *data, uint32_t data_size,
        uint8_t *pcm, uint32_t pcm_capacity_frames, uint32_t *num_frames)
{
    struct LINNEHeader header;
    LINNEApiResult ret;
    uint32_t decoded = 0, bytes;
    if (dec == NULL || data == NULL || pcm == NULL || num_frames == NULL) return LINNE_APIRESULT_INVALID_ARGUMENT;
    *num_frames = 0;
    if ((ret = LINNEDecoder_DecodeHeader(data, data_size, &header)) != LINNE_APIRESULT_OK) return ret;
    if ((ret = LINNEDecoder_SetHeader(dec, &header)) != LINNE_APIRESULT_OK) return ret;
    if (pcm_capacity_frames < header.num_samples) return LINNE_APIRESULT_INSUFFICIENT_BUFFER;
    bytes = header.bits_per_sample / 8u;
    if (bytes == 0 || bytes > 4u || (header.bits_per_sample % 8u) != 0u) return LINNE_APIRESULT_INVALID_FORMAT;
    /* several devices, or a stream long enough to pipeline its transfers against its kernels on one */
    if ((dec->num_devices > 1u || data_size >= (32u << 20))
        && decode_whole_sharded(dec, data, data_size, NULL, pcm, header.num_samples, &ret, &decoded)) {
        *num_frames = decoded;
        return ret;
    }
    ret = decode_packed_range(dec, data, data_size, LINNE_HEADER_SIZE, pcm, header.num_samples, header.num_samples, 0, &decoded);
    *num_frames = decoded;
    return ret;
}
