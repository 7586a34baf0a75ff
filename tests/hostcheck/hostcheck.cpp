// TEST INFRASTRUCTURE: compiles the product's __host__ __device__ arithmetic headers with g++ so the
// CPU-only test tier can compare the exact kernel arithmetic (kb.cuh, poseidon1.cuh) with the oracle
// before any GPU time is spent.  Not linked into the product library and not a fallback path.
#include <cstdint>
#include <cstring>
#include "../../leanmultisig_b200/csrc/poseidon1.cuh"

static const lm::P1Tables H_TAB =
#include "../../leanmultisig_b200/csrc/poseidon1_tables.inc"
    ;

extern "C" {
void hc_poseidon1_permute(uint32_t* states, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) lm::p1_permute<16>(states + 16 * i, H_TAB);
}
void hc_poseidon1_compress8(uint32_t* states, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) lm::p1_compress<8>(states + 16 * i, H_TAB);
}
void hc_ef_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n) {
  for (uint64_t i = 0; i < n; i++) {
    lm::Ef x, y;
    memcpy(&x, a + 5 * i, 20);
    memcpy(&y, b + 5 * i, 20);
    lm::Ef z = lm::ef_mul(x, y);
    memcpy(out + 5 * i, &z, 20);
  }
}
uint32_t hc_kb_mul(uint32_t a, uint32_t b) { return lm::kb_mul(a, b); }
uint32_t hc_kb_add(uint32_t a, uint32_t b) { return lm::kb_add(a, b); }
uint32_t hc_kb_sub(uint32_t a, uint32_t b) { return lm::kb_sub(a, b); }
uint32_t hc_r2() { return lm::KB_R2; }
}
