// TractOracle-Net scoring (oracles/oracle.py:39-89, transformer_oracle.py:77-92).
// Placeholder translation unit: the entry points exist so the ABI is complete; the kernels
// land in a later milestone.  They fail loudly -- there is no CPU fallback.
#include "ttl_common.cuh"

extern "C" {
int ttl_oracle_features(const float*, const int64_t*, int32_t, float*, void*) { return TTL_ERR_UNSUPPORTED; }
int ttl_oracle_forward(const ttl_oracle_weights*, const float*, int32_t, float*, void*) { return TTL_ERR_UNSUPPORTED; }
}
