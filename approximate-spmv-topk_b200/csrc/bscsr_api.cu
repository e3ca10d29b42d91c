// bscsr_api.cu -- placeholder, replaced by the real fixed-point engine.
#include "handle.hpp"
namespace tks {
int bscsr_upload(Handle *h, uint32_t, uint32_t, const uint64_t *, const void *const *, const uint32_t *, const uint64_t *) { return h->fail(TKS_ESTATE, "BS-CSR engine not built"); }
int bscsr_set_query(Handle *h, const uint32_t *, const uint32_t *, cudaStream_t) { return h->fail(TKS_ESTATE, "BS-CSR engine not built"); }
int bscsr_launch(Handle *h, cudaStream_t) { return h->fail(TKS_ESTATE, "BS-CSR engine not built"); }
int bscsr_fetch(Handle *h) { return h->fail(TKS_ESTATE, "BS-CSR engine not built"); }
int bscsr_read_result(Handle *h, uint32_t *, uint32_t *, uint32_t, uint32_t *) { return h->fail(TKS_ESTATE, "BS-CSR engine not built"); }
int bscsr_read_partition_results(Handle *h, uint32_t *, uint32_t *) { return h->fail(TKS_ESTATE, "BS-CSR engine not built"); }
void bscsr_destroy(Handle *) {}
}
extern "C" int tks_upload_bscsr(tks_handle *h, uint32_t cols, uint32_t partitions, const uint64_t *ppp, const void *const *packets, const uint32_t *first_row, const uint64_t *npp) {
    if (!h) return TKS_EINVAL;
    return tks::bscsr_upload(h, cols, partitions, ppp, packets, first_row, npp);
}
