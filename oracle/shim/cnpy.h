// TEST INFRASTRUCTURE ONLY (oracle build shim) -- never included by the product.
//
// Stand-in for rogersce/cnpy (un-vendored dependency of
// /root/reference/graphlily/io/data_loader.h:7).  Only the surface the
// reference loader touches is provided; file parsing is delegated to the
// in-repo npz reader so that the reference's own
// load_csr_matrix_from_float_npz() can be executed unmodified.
#ifndef ORACLE_SHIM_CNPY_H_
#define ORACLE_SHIM_CNPY_H_

#include <cassert>
#include <map>
#include <string>
#include <vector>

#include "../../include/graphlily/io/npz.h"

namespace cnpy {

struct NpyArray {
    std::vector<size_t> shape;
    size_t word_size = 0;
    std::vector<unsigned char> bytes;
    template <typename T> T *data() { return reinterpret_cast<T *>(bytes.data()); }
    template <typename T> const T *data() const { return reinterpret_cast<const T *>(bytes.data()); }
};

typedef std::map<std::string, NpyArray> npz_t;

inline npz_t npz_load(const std::string &path) {
    npz_t out;
    graphlily::io::npz::Archive ar = graphlily::io::npz::load(path);
    for (auto &kv : ar) {
        NpyArray a;
        a.shape = kv.second.shape;
        a.word_size = kv.second.word_size;
        a.bytes.swap(kv.second.bytes);
        out[kv.first] = std::move(a);
    }
    return out;
}

}  // namespace cnpy

#endif  // ORACLE_SHIM_CNPY_H_
