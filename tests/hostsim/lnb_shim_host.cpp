/* lnb_shim_host.cpp -- TEST INFRASTRUCTURE: a loop-based stand-in for the CUDA shim.
 *
 * Implements lnb_shim.h with malloc/memcpy and a sequential executor so the CPU-only CI
 * (pytest -m "not gpu") can exercise the host-side C code and the very same kernel bodies
 * (lnb_*_core.cuh / lnb_pipeline.cuh, compiled here as plain host C++).  It is built into
 * tests/hostsim/liblinne_hostsim.so only; liblinne_b200.so never contains or loads it.
 */
#include <stdlib.h>
#include <string.h>
#include "lnb_shim.h"
#include "lnb_pipeline.cuh"
#include "lnb_hop.cuh"

struct LnbDevice { LnbDevTables tables; uint64_t launches; };

struct LoopExec {
    static constexpr bool cooperative = false;
    LnbDevice *dev;
    void crc_cooperative(const LnbDecodeBatch &) {}
    void tput_decode(const LnbDecodeBatch &) {}
    void entropy_cooperative(const LnbDecodeBatch &) {}
    void synth_cooperative(const LnbDecodeBatch &) {}
    void stream_cooperative(const LnbDecodeBatch &) {}
    uint32_t synth_max_n() const { return 0; }
    template <class F> void run_per_warp(const char *n, uint32_t c, const F &f) { run(n, c, f); }
    template <class F> void run(const char *, uint32_t n, const F &f)
    {
        for (uint32_t i = 0; i < n; i++) f(i);
        dev->launches++;
    }
    void analyze_cooperative(const LnbEncodeBatch &) {}       /* CUDA only; the host never flags blocks for it */
    void scan_cooperative(const LnbEncodeBatch &) {}
    void pack_cooperative(const LnbEncodeBatch &, uint32_t) {}
    void prepare_cooperative(const LnbEncodeBatch &) {}
    void refine_cooperative(const LnbEncodeBatch &, uint32_t) {}
    void predict_plan_cooperative(const LnbEncodeBatch &) {}
};

extern "C" {
const char *lnb_shim_backend(void) { return "hostsim"; }
int lnb_shim_device_count(void) { return 1; }
int lnb_shim_set_device(int) { return 0; }
int lnb_shim_current_device(void) { return 0; }
int lnb_shim_device_ordinal(const LnbDevice *) { return 0; }
uint32_t lnb_shim_fast_max_na(void) { return 0; }
uint32_t lnb_shim_coop_max_n(void) { return 0; }
uint32_t lnb_shim_refine_max_na(void) { return 0; }
uint32_t lnb_shim_refine_hist(void) { return 0; }
uint32_t lnb_shim_fused_max_n(void) { return 0; }
int lnb_shim_tput_supported(const LnbStreamCfg *) { return 0; }
int lnb_shim_open(LnbDevice **out, int)
{
    LnbDevice *dev = (LnbDevice *)calloc(1, sizeof(LnbDevice));
    const LnbHostTables *ht = lnb_tables_get();
    dev->tables.huff_lut = ht->huff_lut; dev->tables.huff_code = ht->huff_code; dev->tables.huff_len = ht->huff_len;
    dev->tables.k2_threshold = ht->k2_threshold; dev->tables.crc_table = ht->crc_table;
    *out = dev;
    return 0;
}
void lnb_shim_close(LnbDevice *dev) { free(dev); }
const LnbDevTables *lnb_shim_tables(const LnbDevice *dev) { return &dev->tables; }
void lnb_shim_use_stream(LnbDevice *, void *) {}
void lnb_shim_set_cost_rank(LnbDevice *, int) {}
void *lnb_shim_alloc(LnbDevice *, size_t bytes) { return calloc(1, bytes ? bytes : 16); }
void lnb_shim_free(LnbDevice *, void *p) { free(p); }
void *lnb_shim_alloc_pinned(size_t bytes) { return calloc(1, bytes ? bytes : 16); }
void lnb_shim_free_pinned(void *p) { free(p); }
int lnb_shim_h2d(LnbDevice *, void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); return 0; }
int lnb_shim_d2h(LnbDevice *, void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); return 0; }
int lnb_shim_d2d(LnbDevice *, void *d, const void *s, size_t n) { if (n) memmove(d, s, n); return 0; }
int lnb_shim_memset(LnbDevice *, void *d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
int lnb_shim_sync(LnbDevice *) { return 0; }
int lnb_shim_decode(LnbDevice *dev, const LnbDecodeBatch *b) { LoopExec ex{dev}; lnb_decode_pipeline(ex, *b); return 0; }
int lnb_shim_encode_analyze(LnbDevice *dev, const LnbEncodeBatch *b) { LoopExec ex{dev}; lnb_encode_analyze_pipeline(ex, *b); return 0; }
int lnb_shim_encode_pack(LnbDevice *dev, const LnbEncodeBatch *b, uint32_t cap) { LoopExec ex{dev}; lnb_encode_pack_pipeline(ex, *b, cap); return 0; }
int lnb_shim_hop(LnbDevice *dev, const uint8_t *img, const LnbHopFile *files, uint32_t n, LnbBlockDesc *table, LnbHopResult *res)
{ for (uint32_t i = 0; i < n; i++) lnb_hop_file(img, files[i], table + files[i].table_first, res[i]); dev->launches++; return 0; }
int lnb_shim_unpack_pcm(LnbDevice *dev, const uint8_t *pk, int32_t *pcm, uint32_t stride, uint32_t frames, uint32_t ch, uint32_t bytes)
{ LoopExec ex{dev}; lnb_unpack_pcm_pipeline(ex, pk, pcm, stride, frames, ch, bytes); return 0; }
int lnb_shim_pack_pcm(LnbDevice *dev, const int32_t *pcm, uint8_t *pk, uint32_t stride, uint32_t frames, uint32_t ch, uint32_t bytes)
{ LoopExec ex{dev}; lnb_pack_pcm_pipeline(ex, pcm, pk, stride, frames, ch, bytes); return 0; }
uint64_t lnb_shim_launch_count(const LnbDevice *dev) { return dev->launches; }
void lnb_shim_profile_enable(LnbDevice *, int) {}
void lnb_shim_profile_reset(LnbDevice *) {}
int lnb_shim_profile_get(LnbDevice *, LnbStageStat *, int) { return 0; }
int lnb_shim_profile_timeline(LnbDevice *, LnbTimelineEntry *, int) { return 0; }
double lnb_shim_measure_fp64_tflops(LnbDevice *) { return 0.0; }
int lnb_shim_fill_sine_window(LnbDevice *, double *, uint32_t) { return 1; }
void *lnb_shim_device_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 16); }
void lnb_shim_device_free(void *p) { free(p); }
int lnb_shim_ipc_export(const void *, unsigned char *) { return 1; }          /* no peers on the CPU */
void *lnb_shim_ipc_open(const unsigned char *) { return NULL; }
void lnb_shim_ipc_close(void *) {}
int lnb_shim_copy(void *d, const void *s, size_t n, int) { if (n) memcpy(d, s, n); return 0; }
}
