// graphlily-b200: CSR / CSC containers, scipy .npz loader, transpose.
//
// Same types and functions as /root/reference/graphlily/io/data_loader.h: CSRMatrix / CSCMatrix
// (:19-31,92-104), create_csr_matrix (:35-47), load_csr_matrix_from_float_npz (:51-70),
// csr_matrix_convert_from_float (:75-84), csr2csc (:108-144), csc_matrix_convert_from_float
// (:148-157).  cnpy is replaced by the in-repo reader graphlily/io/npz.h.
#ifndef GRAPHLILY_IO_DATA_LOADER_H_
#define GRAPHLILY_IO_DATA_LOADER_H_

#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "graphlily/io/npz.h"

namespace graphlily {
namespace io {

template <typename data_type>
struct CSRMatrix {
    uint32_t num_rows;
    uint32_t num_cols;
    std::vector<data_type> adj_data;
    std::vector<uint32_t> adj_indices;   // column indices
    std::vector<uint32_t> adj_indptr;    // num_rows + 1
};

template <typename data_type>
CSRMatrix<data_type> create_csr_matrix(uint32_t num_rows, uint32_t num_cols, std::vector<data_type> const &adj_data,
                                       std::vector<uint32_t> const &adj_indices,
                                       std::vector<uint32_t> const &adj_indptr) {
    return CSRMatrix<data_type>{num_rows, num_cols, adj_data, adj_indices, adj_indptr};
}

namespace detail {
// scipy writes int32 index arrays (int64 for huge matrices) and an int64 shape; the reference
// reinterprets them as uint32 words (data_loader.h:55-56).  Accept either width.
inline std::vector<uint32_t> as_u32(const npz::Array &a, const char *name) {
    const size_t n = a.num_elements();
    std::vector<uint32_t> out(n);
    if (a.word_size == 4) {
        const uint32_t *p = a.data<uint32_t>();
        out.assign(p, p + n);
    } else if (a.word_size == 8) {
        const uint64_t *p = a.data<uint64_t>();
        for (size_t i = 0; i < n; i++) out[i] = uint32_t(p[i]);
    } else {
        throw std::runtime_error(std::string("npz: unsupported integer width in ") + name);
    }
    return out;
}
}  // namespace detail

// Load a csr matrix from a scipy sparse npz file. The sparse matrix should have float data type.
// The reference's tests and run_*.sh scripts name datasets by absolute paths of its authors' machine
// (/work/shared/common/project_build/graphblas/data/sparse_matrix_graph/<name>.npz).  When a path does
// not exist and GLB_DATASET_DIR is set, the file of the same name in that directory is loaded instead,
// so those callers run unchanged.
inline CSRMatrix<float> load_csr_matrix_from_float_npz(std::string csr_float_npz_path) {
    if (const char *dir = std::getenv("GLB_DATASET_DIR")) {
        std::ifstream probe(csr_float_npz_path, std::ios::binary);
        if (!probe) {
            const size_t slash = csr_float_npz_path.find_last_of('/');
            csr_float_npz_path = std::string(dir) + "/" +
                                 (slash == std::string::npos ? csr_float_npz_path : csr_float_npz_path.substr(slash + 1));
        }
    }
    npz::Archive z = npz::load(csr_float_npz_path);
    for (const char *k : {"shape", "data", "indices", "indptr"})
        if (!z.count(k)) throw std::runtime_error(std::string("npz: missing member ") + k);
    CSRMatrix<float> m;
    std::vector<uint32_t> shape = detail::as_u32(z["shape"], "shape");
    if (shape.size() != 2) throw std::runtime_error("npz: shape must have two entries");
    m.num_rows = shape[0];
    m.num_cols = shape[1];
    const npz::Array &d = z["data"];
    if (d.descr != "<f4") throw std::runtime_error("npz: data must be float32 (got " + d.descr + ")");
    m.adj_data.assign(d.data<float>(), d.data<float>() + d.num_elements());
    m.adj_indices = detail::as_u32(z["indices"], "indices");
    m.adj_indptr = detail::as_u32(z["indptr"], "indptr");
    if (m.adj_indptr.size() != size_t(m.num_rows) + 1 || m.adj_indices.size() != m.adj_data.size())
        throw std::runtime_error("npz: inconsistent CSR arrays");
    return m;
}

template <typename data_type>
CSRMatrix<data_type> csr_matrix_convert_from_float(CSRMatrix<float> const &in) {
    CSRMatrix<data_type> out;
    out.num_rows = in.num_rows;
    out.num_cols = in.num_cols;
    out.adj_data.assign(in.adj_data.begin(), in.adj_data.end());
    out.adj_indices = in.adj_indices;
    out.adj_indptr = in.adj_indptr;
    return out;
}

template <typename data_type>
struct CSCMatrix {
    uint32_t num_rows;
    uint32_t num_cols;
    std::vector<data_type> adj_data;
    std::vector<uint32_t> adj_indices;   // row indices
    std::vector<uint32_t> adj_indptr;    // num_cols + 1
};

// Counting-sort transpose; inside a column the row order of the CSR is preserved.
template <typename data_type>
CSCMatrix<data_type> csr2csc(CSRMatrix<data_type> const &csr) {
    CSCMatrix<data_type> csc;
    csc.num_rows = csr.num_rows;
    csc.num_cols = csr.num_cols;
    const size_t nnz = csr.adj_indptr[csr.num_rows];
    csc.adj_data.resize(nnz);
    csc.adj_indices.resize(nnz);
    csc.adj_indptr.assign(size_t(csc.num_cols) + 1, 0);
    for (size_t i = 0; i < nnz; i++) csc.adj_indptr[csr.adj_indices[i] + 1]++;
    for (size_t c = 0; c < csc.num_cols; c++) csc.adj_indptr[c + 1] += csc.adj_indptr[c];
    assert(csc.adj_indptr[csc.num_cols] == nnz);
    std::vector<uint32_t> cursor(csc.adj_indptr.begin(), csc.adj_indptr.end() - 1);
    for (uint32_t r = 0; r < csr.num_rows; r++) {
        for (size_t i = csr.adj_indptr[r]; i < csr.adj_indptr[r + 1]; i++) {
            const uint32_t dest = cursor[csr.adj_indices[i]]++;
            csc.adj_indices[dest] = r;
            csc.adj_data[dest] = csr.adj_data[i];
        }
    }
    return csc;
}

template <typename data_type>
CSCMatrix<data_type> csc_matrix_convert_from_float(CSCMatrix<float> const &in) {
    CSCMatrix<data_type> out;
    out.num_rows = in.num_rows;
    out.num_cols = in.num_cols;
    out.adj_data.assign(in.adj_data.begin(), in.adj_data.end());
    out.adj_indices = in.adj_indices;
    out.adj_indptr = in.adj_indptr;
    return out;
}

}  // namespace io
}  // namespace graphlily

#endif  // GRAPHLILY_IO_DATA_LOADER_H_
