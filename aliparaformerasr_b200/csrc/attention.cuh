// SAN-M / cross attention launcher. See attention.cu.
#pragma once
#include "common.cuh"

namespace pf {

// Q: rows (b*Tq + i), K/V: rows (b*Tk + j); head h occupies columns [h*head_dim, (h+1)*head_dim) of each row.
void attention_launch(const __half* Q, const __half* K, const __half* V, __half* O, int B, int H, int Tq, int Tk, int ldq,
                      int ldk, int ldv, int ldo, int head_dim, cudaStream_t s);

}  // namespace pf
