/* linne_b200_ext.c -- extension entry points declared in include/linne_b200.h. */
#include "linne_b200.h"
#include "lnb_shim.h"

/* both handle structs start with their LINNEHeader; the device context is reached through accessors
 * defined next to the struct definitions */
LnbDevice *lnb_encoder_device(const struct LINNEEncoder *enc);
LnbDevice *lnb_decoder_device(const struct LINNEDecoder *dec);
void lnb_decoder_set_readahead(struct LINNEDecoder *dec, uint32_t blocks);
void lnb_decoder_set_tput_min_blocks(struct LINNEDecoder *dec, uint32_t blocks);

const char *LINNEB200_Backend(void) { return lnb_shim_backend(); }

int LINNEB200_DeviceAvailable(void)
{
    LnbDevice *dev = NULL;
    if (lnb_shim_open(&dev, -1) != 0) return 0;
    lnb_shim_close(dev);
    return 1;
}

uint64_t LINNEB200_EncoderLaunchCount(const struct LINNEEncoder *e) { return e ? lnb_shim_launch_count(lnb_encoder_device(e)) : 0; }
uint64_t LINNEB200_DecoderLaunchCount(const struct LINNEDecoder *d) { return d ? lnb_shim_launch_count(lnb_decoder_device(d)) : 0; }
void LINNEB200_EncoderUseStream(struct LINNEEncoder *e, void *s) { if (e) lnb_shim_use_stream(lnb_encoder_device(e), s); }
void LINNEB200_DecoderUseStream(struct LINNEDecoder *d, void *s) { if (d) lnb_shim_use_stream(lnb_decoder_device(d), s); }

void LINNEB200_DecoderSetReadahead(struct LINNEDecoder *d, uint32_t blocks) { if (d) lnb_decoder_set_readahead(d, blocks); }
void LINNEB200_DecoderSetThroughputBlocks(struct LINNEDecoder *d, uint32_t min_blocks) { if (d) lnb_decoder_set_tput_min_blocks(d, min_blocks); }
void *LINNEB200_HostAlloc(size_t bytes) { return lnb_shim_alloc_pinned(bytes); }
void LINNEB200_HostFree(void *h_ptr) { if (h_ptr) lnb_shim_free_pinned(h_ptr); }

void LINNEB200_EncoderSetProfiling(struct LINNEEncoder *e, int on) { if (e) lnb_shim_profile_enable(lnb_encoder_device(e), on); }
void LINNEB200_DecoderSetProfiling(struct LINNEDecoder *d, int on) { if (d) lnb_shim_profile_enable(lnb_decoder_device(d), on); }
void LINNEB200_EncoderResetStageStats(struct LINNEEncoder *e) { if (e) lnb_shim_profile_reset(lnb_encoder_device(e)); }
void LINNEB200_DecoderResetStageStats(struct LINNEDecoder *d) { if (d) lnb_shim_profile_reset(lnb_decoder_device(d)); }
int LINNEB200_EncoderGetStageStats(struct LINNEEncoder *e, struct LINNEB200StageStat *out, int n)
{
    return e ? lnb_shim_profile_get(lnb_encoder_device(e), (LnbStageStat *)out, n) : 0;
}
int LINNEB200_DecoderGetStageStats(struct LINNEDecoder *d, struct LINNEB200StageStat *out, int n)
{
    return d ? lnb_shim_profile_get(lnb_decoder_device(d), (LnbStageStat *)out, n) : 0;
}
double LINNEB200_MeasureFp64Tflops(void)
{
    LnbDevice *dev = NULL;
    double t;
    if (lnb_shim_open(&dev, -1) != 0) return 0.0;
    t = lnb_shim_measure_fp64_tflops(dev);
    lnb_shim_close(dev);
    return t;
}

/* ---- device memory and peer mappings (multi-GPU sharding) ---- */
void *LINNEB200_DeviceAlloc(size_t bytes) { return lnb_shim_device_alloc(bytes); }
void LINNEB200_DeviceFree(void *d_ptr) { lnb_shim_device_free(d_ptr); }
int LINNEB200_IpcExport(const void *d_ptr, uint8_t handle[64]) { return (d_ptr && handle) ? lnb_shim_ipc_export(d_ptr, handle) : 1; }
void *LINNEB200_IpcOpen(const uint8_t handle[64]) { return handle ? lnb_shim_ipc_open(handle) : NULL; }
void LINNEB200_IpcClose(void *d_peer_ptr) { lnb_shim_ipc_close(d_peer_ptr); }
int LINNEB200_DeviceCopy(void *d_dst, const void *d_src, size_t bytes) { return lnb_shim_copy(d_dst, d_src, bytes, 0); }
int LINNEB200_CopyToDevice(void *d_dst, const void *h_src, size_t bytes) { return lnb_shim_copy(d_dst, h_src, bytes, 1); }
int LINNEB200_CopyToHost(void *h_dst, const void *d_src, size_t bytes) { return lnb_shim_copy(h_dst, d_src, bytes, 2); }

int LINNEB200_EncoderGetTimeline(struct LINNEEncoder *e, struct LINNEB200TimelineEntry *out, int n)
{
    return e ? lnb_shim_profile_timeline(lnb_encoder_device(e), (LnbTimelineEntry *)out, n) : 0;
}
int LINNEB200_DecoderGetTimeline(struct LINNEDecoder *d, struct LINNEB200TimelineEntry *out, int n)
{
    return d ? lnb_shim_profile_timeline(lnb_decoder_device(d), (LnbTimelineEntry *)out, n) : 0;
}
