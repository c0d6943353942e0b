/* lnb_shim_cuda.cu -- CUDA implementation of lnb_shim.h for sm_100a (B200).
 *
 * Round-1 kernel shape: every pipeline stage is a flat grid of independent work items (one thread
 * per block / block-channel / unit, see lnb_pipeline.cuh).  The grid is sized in whole waves of
 * 148 SMs where the item count allows.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lnb_shim.h"
#include "lnb_pipeline.cuh"

struct LnbDevice {
    cudaStream_t stream;
    int owns_stream;
    int ordinal;
    LnbDevTables tables;
    void *table_mem;
    uint64_t launches;
    cudaError_t last_error;
};

template <class F>
__global__ void __launch_bounds__(128) lnb_items_kernel(uint32_t n, F f)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

struct CudaExec {
    LnbDevice *dev;
    template <class F> void run(const char *, uint32_t n, const F &f)
    {
        if (n == 0) return;
        const uint32_t threads = 128;
        const uint32_t grid = (n + threads - 1) / threads;
        lnb_items_kernel<F><<<grid, threads, 0, dev->stream>>>(n, f);
        dev->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && dev->last_error == cudaSuccess) dev->last_error = e;
    }
};

extern "C" {

const char *lnb_shim_backend(void) { return "cuda-sm_100a"; }

int lnb_shim_open(LnbDevice **out, int device_ordinal)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return 1;
    if (device_ordinal >= 0) {
        if (device_ordinal >= count || cudaSetDevice(device_ordinal) != cudaSuccess) return 2;
    }
    LnbDevice *dev = (LnbDevice *)calloc(1, sizeof(LnbDevice));
    if (!dev) return 3;
    cudaGetDevice(&dev->ordinal);
    if (cudaStreamCreateWithFlags(&dev->stream, cudaStreamNonBlocking) != cudaSuccess) { free(dev); return 4; }
    dev->owns_stream = 1;

    const LnbHostTables *ht = lnb_tables_get();
    const size_t sz_lut = sizeof(ht->huff_lut), sz_code = sizeof(ht->huff_code), sz_len = 256,
                 sz_thr = sizeof(ht->k2_threshold), sz_crc = sizeof(ht->crc_table);
    const size_t total = sz_lut + sz_code + sz_thr + sz_crc + sz_len + 64;
    uint8_t *mem = NULL;
    if (cudaMalloc((void **)&mem, total) != cudaSuccess) { cudaStreamDestroy(dev->stream); free(dev); return 5; }
    size_t off = 0;
    /* doubles first (8-byte alignment), then 4-, 2-, 1-byte tables */
    cudaMemcpy(mem + off, ht->k2_threshold, sz_thr, cudaMemcpyHostToDevice); dev->tables.k2_threshold = (const double *)(mem + off); off += sz_thr;
    cudaMemcpy(mem + off, ht->huff_code, sz_code, cudaMemcpyHostToDevice);   dev->tables.huff_code = (const uint32_t *)(mem + off); off += sz_code;
    cudaMemcpy(mem + off, ht->huff_lut, sz_lut, cudaMemcpyHostToDevice);     dev->tables.huff_lut = (const uint16_t *)(mem + off); off += sz_lut;
    cudaMemcpy(mem + off, ht->crc_table, sz_crc, cudaMemcpyHostToDevice);    dev->tables.crc_table = (const uint16_t *)(mem + off); off += sz_crc;
    cudaMemcpy(mem + off, ht->huff_len, sz_len, cudaMemcpyHostToDevice);     dev->tables.huff_len = (const uint8_t *)(mem + off); off += sz_len;
    dev->table_mem = mem;
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(mem); cudaStreamDestroy(dev->stream); free(dev); return 6; }
    *out = dev;
    return 0;
}

void lnb_shim_close(LnbDevice *dev)
{
    if (!dev) return;
    cudaStreamSynchronize(dev->stream);
    cudaFree(dev->table_mem);
    if (dev->owns_stream) cudaStreamDestroy(dev->stream);
    free(dev);
}

const LnbDevTables *lnb_shim_tables(const LnbDevice *dev) { return &dev->tables; }

void lnb_shim_use_stream(LnbDevice *dev, void *cuda_stream)
{
    if (dev->owns_stream) { cudaStreamSynchronize(dev->stream); cudaStreamDestroy(dev->stream); dev->owns_stream = 0; }
    dev->stream = (cudaStream_t)cuda_stream;
}

void *lnb_shim_alloc(LnbDevice *dev, size_t bytes)
{
    void *p = NULL;
    (void)dev;
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return p;
}
void lnb_shim_free(LnbDevice *dev, void *ptr) { (void)dev; if (ptr) cudaFree(ptr); }
void *lnb_shim_alloc_pinned(size_t bytes)
{
    void *p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return p;
}
void lnb_shim_free_pinned(void *ptr) { if (ptr) cudaFreeHost(ptr); }

static int note(LnbDevice *dev, cudaError_t e)
{
    if (e != cudaSuccess && dev->last_error == cudaSuccess) dev->last_error = e;
    return e == cudaSuccess ? 0 : 1;
}
int lnb_shim_h2d(LnbDevice *dev, void *dst, const void *src, size_t bytes)
{
    return bytes ? note(dev, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, dev->stream)) : 0;
}
int lnb_shim_d2h(LnbDevice *dev, void *dst, const void *src, size_t bytes)
{
    return bytes ? note(dev, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream)) : 0;
}
int lnb_shim_memset(LnbDevice *dev, void *dst, int value, size_t bytes)
{
    return bytes ? note(dev, cudaMemsetAsync(dst, value, bytes, dev->stream)) : 0;
}
int lnb_shim_sync(LnbDevice *dev)
{
    cudaError_t e = cudaStreamSynchronize(dev->stream);
    if (e == cudaSuccess) e = dev->last_error;
    if (e != cudaSuccess) {
        fprintf(stderr, "linne_b200: CUDA error: %s\n", cudaGetErrorString(e));
        dev->last_error = cudaSuccess;
        cudaGetLastError();
        return 1;
    }
    return 0;
}

int lnb_shim_decode(LnbDevice *dev, const LnbDecodeBatch *batch)
{
    CudaExec ex{dev};
    lnb_decode_pipeline(ex, *batch);
    return dev->last_error == cudaSuccess ? 0 : 1;
}
int lnb_shim_encode_analyze(LnbDevice *dev, const LnbEncodeBatch *batch)
{
    CudaExec ex{dev};
    lnb_encode_analyze_pipeline(ex, *batch);
    return dev->last_error == cudaSuccess ? 0 : 1;
}
int lnb_shim_encode_pack(LnbDevice *dev, const LnbEncodeBatch *batch, uint32_t out_capacity)
{
    CudaExec ex{dev};
    lnb_encode_pack_pipeline(ex, *batch, out_capacity);
    return dev->last_error == cudaSuccess ? 0 : 1;
}

uint64_t lnb_shim_launch_count(const LnbDevice *dev) { return dev->launches; }

}  /* extern "C" */
