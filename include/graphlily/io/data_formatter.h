// graphlily-b200: matrix pre-processing helpers the apps run before upload.
//
// /root/reference/graphlily/io/data_formatter.h:19-51: util_round_csr_matrix_dim and
// util_normalize_csr_matrix_by_outdegree.  The rest of that file (csr2cpsr :457-534, formatCSC
// :608-721) builds the FPGA's packet layouts; the device layout of this engine is built inside
// glb_csr_create / glb_csc_create (graphlily_b200/csrc/spmv.cu, spmspv.cu).
#ifndef GRAPHLILY_IO_DATA_FORMATTER_H_
#define GRAPHLILY_IO_DATA_FORMATTER_H_

#include <cstdint>
#include <vector>

#include "graphlily/global.h"
#include "graphlily/io/data_loader.h"

namespace graphlily {
namespace io {

// Pad rows (repeating the last indptr entry) and columns up to multiples of the divisors.
template <typename data_type>
void util_round_csr_matrix_dim(CSRMatrix<data_type> &csr_matrix, uint32_t row_divisor, uint32_t col_divisor) {
    if (csr_matrix.num_rows % row_divisor != 0) {
        const uint32_t pad = row_divisor - csr_matrix.num_rows % row_divisor;
        csr_matrix.adj_indptr.insert(csr_matrix.adj_indptr.end(), pad, csr_matrix.adj_indptr[csr_matrix.num_rows]);
        csr_matrix.num_rows += pad;
    }
    if (csr_matrix.num_cols % col_divisor != 0) csr_matrix.num_cols += col_divisor - csr_matrix.num_cols % col_divisor;
}

// data[i] = 1.0 / (number of non-zeros in the column of i): double divide, then the store narrows.
template <typename data_type>
void util_normalize_csr_matrix_by_outdegree(CSRMatrix<data_type> &csr_matrix) {
    std::vector<uint32_t> nnz_each_col(csr_matrix.num_cols, 0);
    for (auto col_idx : csr_matrix.adj_indices) nnz_each_col[col_idx]++;
    const size_t nnz = csr_matrix.adj_indptr[csr_matrix.num_rows];
    for (size_t i = 0; i < nnz; i++) csr_matrix.adj_data[i] = 1.0 / nnz_each_col[csr_matrix.adj_indices[i]];
}

}  // namespace io
}  // namespace graphlily

#endif  // GRAPHLILY_IO_DATA_FORMATTER_H_
