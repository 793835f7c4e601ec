// graphlily-b200: minimal .npz (zip of .npy members) reader / writer.
//
// Replaces the cnpy dependency of the reference loader
// (/root/reference/graphlily/io/data_loader.h:7,51-70: cnpy::npz_load on a
// scipy.sparse.save_npz file).  scipy writes deflate-compressed members
// `indices`/`indptr` (int32), `data` (float32), `shape` (int64[2]) and
// `format`; numpy's savez writes them with zip64 local headers, so member
// sizes are taken from the central directory.
//
// Reader: zip central directory walk + zlib raw inflate (method 8) or plain
// copy (method 0) + .npy v1/v2/v3 header parse.  Writer: stored (method 0)
// members with CRC32, enough for numpy.load / scipy.sparse.load_npz.
#ifndef GRAPHLILY_B200_IO_NPZ_H_
#define GRAPHLILY_B200_IO_NPZ_H_

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace graphlily {
namespace io {
namespace npz {

/*! \brief One decoded .npy array: raw little-endian payload + dtype descriptor. */
struct Array {
    std::vector<size_t> shape;
    std::string descr;      // e.g. "<f4", "<i4", "<i8", "|S3"
    size_t word_size = 0;   // bytes per element
    bool fortran_order = false;
    std::vector<unsigned char> bytes;

    size_t num_elements() const {
        size_t n = 1;
        for (size_t d : shape) n *= d;
        return n;
    }
    template <typename T> T *data() { return reinterpret_cast<T *>(bytes.data()); }
    template <typename T> const T *data() const { return reinterpret_cast<const T *>(bytes.data()); }
};

typedef std::map<std::string, Array> Archive;

namespace detail {

inline uint16_t rd16(const unsigned char *p) { return uint16_t(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const unsigned char *p) {
    return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}
inline uint64_t rd64(const unsigned char *p) { return uint64_t(rd32(p)) | (uint64_t(rd32(p + 4)) << 32); }

inline void fail(const std::string &what) { throw std::runtime_error("npz: " + what); }

// Parse the python-dict header of a .npy member.
inline void parse_npy(const std::vector<unsigned char> &raw, Array &out) {
    if (raw.size() < 10 || raw[0] != 0x93 || std::memcmp(&raw[1], "NUMPY", 5) != 0) fail("bad .npy magic");
    unsigned major = raw[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = rd16(&raw[8]); hoff = 10; }
    else { if (raw.size() < 12) fail("short .npy"); hlen = rd32(&raw[8]); hoff = 12; }
    if (hoff + hlen > raw.size()) fail("truncated .npy header");
    std::string hdr(reinterpret_cast<const char *>(&raw[hoff]), hlen);

    size_t p = hdr.find("'descr'");
    if (p == std::string::npos) fail("no descr");
    p = hdr.find(':', p);
    if (p != std::string::npos) p = hdr.find('\'', p);
    size_t q = (p == std::string::npos) ? p : hdr.find('\'', p + 1);
    if (p == std::string::npos || q == std::string::npos) fail("malformed descr");
    out.descr = hdr.substr(p + 1, q - p - 1);
    if (out.descr.size() < 3) fail("unsupported descr " + out.descr);
    if (out.descr[0] == '>') fail("big-endian arrays unsupported");
    for (size_t k = 2; k < out.descr.size(); k++)
        if (out.descr[k] < '0' || out.descr[k] > '9') fail("unsupported descr " + out.descr);
    out.word_size = size_t(std::stoul(out.descr.substr(2)));
    if (out.descr[1] == 'U') out.word_size *= 4;

    p = hdr.find("'fortran_order'");
    if (p == std::string::npos) fail("no fortran_order");
    p = hdr.find_first_not_of(" :", p + 15);
    if (p == std::string::npos) fail("malformed fortran_order");
    out.fortran_order = hdr.compare(p, 4, "True") == 0;

    p = hdr.find("'shape'");
    if (p == std::string::npos) fail("no shape");
    p = hdr.find('(', p);
    q = (p == std::string::npos) ? p : hdr.find(')', p);
    if (p == std::string::npos || q == std::string::npos) fail("malformed shape");
    out.shape.clear();
    std::string dims = hdr.substr(p + 1, q - p - 1);
    size_t i = 0;
    while (i < dims.size()) {
        while (i < dims.size() && (dims[i] < '0' || dims[i] > '9')) i++;
        if (i >= dims.size()) break;
        size_t v = 0;
        while (i < dims.size() && dims[i] >= '0' && dims[i] <= '9') v = v * 10 + size_t(dims[i++] - '0');
        out.shape.push_back(v);
    }
    size_t payload = 0, limit = raw.size() - hoff - hlen;
    {   // the product must not wrap: every partial product is checked against what the member can hold
        size_t n = out.word_size ? 1 : 0;
        for (size_t d : out.shape) {
            if (d != 0 && n > limit / d) fail("shape larger than the .npy payload");
            n *= d;
        }
        if (out.word_size != 0 && n > limit / out.word_size) fail("truncated .npy payload");
        payload = n * out.word_size;
    }
    out.bytes.assign(raw.begin() + long(hoff + hlen), raw.begin() + long(hoff + hlen + payload));
}

inline std::vector<unsigned char> inflate_raw(const unsigned char *src, size_t csize, size_t usize) {
    std::vector<unsigned char> out(usize ? usize : 1);
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -MAX_WBITS) != Z_OK) fail("inflateInit2 failed");
    size_t in_done = 0, out_done = 0;
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
        size_t in_chunk = std::min<size_t>(csize - in_done, size_t(1) << 30);
        size_t out_chunk = std::min<size_t>(usize - out_done, size_t(1) << 30);
        zs.next_in = const_cast<unsigned char *>(src + in_done);
        zs.avail_in = uInt(in_chunk);
        zs.next_out = out.data() + out_done;
        zs.avail_out = uInt(out_chunk);
        rc = inflate(&zs, Z_NO_FLUSH);
        in_done += in_chunk - zs.avail_in;
        out_done += out_chunk - zs.avail_out;
        if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); fail("inflate error"); }
        if (rc == Z_OK && in_chunk == 0 && out_chunk == 0) { inflateEnd(&zs); fail("inflate stalled"); }
    }
    inflateEnd(&zs);
    if (out_done != usize) fail("inflate size mismatch");
    out.resize(usize);
    return out;
}

}  // namespace detail

/*! \brief Load every member of an .npz archive. Member names drop the ".npy" suffix. */
inline Archive load(const std::string &path) {
    using namespace detail;
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) fail("cannot open " + path);
    size_t fsize = size_t(f.tellg());
    std::vector<unsigned char> buf(fsize);
    f.seekg(0);
    f.read(reinterpret_cast<char *>(buf.data()), std::streamsize(fsize));
    if (fsize < 22) fail("file too small");

    // End-of-central-directory record (scan backwards over a possible comment).
    size_t eocd = std::string::npos;
    for (size_t i = fsize - 22;; i--) {
        if (rd32(&buf[i]) == 0x06054b50u) { eocd = i; break; }
        if (i == 0 || fsize - i > 22 + 65535) break;
    }
    if (eocd == std::string::npos) fail("no end-of-central-directory");
    uint64_t n_entries = rd16(&buf[eocd + 10]);
    uint64_t cd_off = rd32(&buf[eocd + 16]);
    if (n_entries == 0xFFFF || cd_off == 0xFFFFFFFFu) {
        // zip64: locator sits right before the EOCD
        if (eocd < 20 || rd32(&buf[eocd - 20]) != 0x07064b50u) fail("zip64 locator missing");
        uint64_t e64 = rd64(&buf[eocd - 20 + 8]);
        if (e64 + 56 > fsize || rd32(&buf[e64]) != 0x06064b50u) fail("zip64 EOCD missing");
        n_entries = rd64(&buf[e64 + 32]);
        cd_off = rd64(&buf[e64 + 48]);
    }

    Archive out;
    size_t p = size_t(cd_off);
    for (uint64_t e = 0; e < n_entries; e++) {
        if (p + 46 > fsize || rd32(&buf[p]) != 0x02014b50u) fail("bad central directory entry");
        uint16_t method = rd16(&buf[p + 10]);
        uint64_t csize = rd32(&buf[p + 20]);
        uint64_t usize = rd32(&buf[p + 24]);
        uint16_t nlen = rd16(&buf[p + 28]), xlen = rd16(&buf[p + 30]), clen = rd16(&buf[p + 32]);
        uint64_t lho = rd32(&buf[p + 42]);
        if (p + 46 + size_t(nlen) + xlen + clen > fsize) fail("central directory entry overruns the file");
        std::string name(reinterpret_cast<const char *>(&buf[p + 46]), nlen);
        // zip64 extended information (header id 0x0001): fields present only for saturated values
        size_t x = p + 46 + nlen, xend = x + xlen;
        while (x + 4 <= xend) {
            uint16_t id = rd16(&buf[x]), sz = rd16(&buf[x + 2]);
            if (x + 4 + size_t(sz) > xend) fail("extra field overruns its entry: " + name);
            if (id == 0x0001) {
                size_t q = x + 4, qend = x + 4 + sz;
                auto take64 = [&](uint64_t &v) {
                    if (q + 8 > qend) fail("truncated zip64 extra field: " + name);
                    v = rd64(&buf[q]);
                    q += 8;
                };
                if (usize == 0xFFFFFFFFu) take64(usize);
                if (csize == 0xFFFFFFFFu) take64(csize);
                if (lho == 0xFFFFFFFFu) take64(lho);
            }
            x += 4 + sz;
        }
        p = xend + clen;

        if (lho > fsize || lho + 30 > fsize || rd32(&buf[lho]) != 0x04034b50u) fail("bad local header for " + name);
        size_t data_off = size_t(lho) + 30 + rd16(&buf[lho + 26]) + rd16(&buf[lho + 28]);
        if (data_off > fsize || csize > fsize - data_off) fail("member overruns file: " + name);

        std::vector<unsigned char> raw;
        if (method == 0) raw.assign(buf.begin() + long(data_off), buf.begin() + long(data_off + csize));
        else if (method == 8) raw = inflate_raw(&buf[data_off], size_t(csize), size_t(usize));
        else fail("unsupported compression method for " + name);

        if (name.size() > 4 && name.compare(name.size() - 4, 4, ".npy") == 0) name.resize(name.size() - 4);
        parse_npy(raw, out[name]);
    }
    return out;
}

/*! \brief Incremental writer of an uncompressed .npz (stored members). */
class Writer {
public:
    explicit Writer(const std::string &path) : f_(path, std::ios::binary) {
        if (!f_) detail::fail("cannot create " + path);
    }

    /*! \brief Append a C-contiguous array. `descr` is the numpy dtype string, e.g. "<f4". */
    void add(const std::string &name, const std::string &descr, const std::vector<size_t> &shape,
             const void *data, size_t nbytes) {
        std::string dims = "(";
        for (size_t i = 0; i < shape.size(); i++) dims += std::to_string(shape[i]) + ",";
        dims += ")";
        std::string dict = "{'descr': '" + descr + "', 'fortran_order': False, 'shape': " + dims + ", }";
        size_t unpadded = 10 + dict.size() + 1;
        dict.append((64 - unpadded % 64) % 64, ' ');
        dict.push_back('\n');
        std::vector<unsigned char> hdr = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0,
                                          (unsigned char)(dict.size() & 0xFF), (unsigned char)(dict.size() >> 8)};
        hdr.insert(hdr.end(), dict.begin(), dict.end());

        uint64_t total = hdr.size() + nbytes;
        if (total >= 0xFFFFFFFFull) detail::fail("member too large for zip32 writer: " + name);
        if (dict.size() > 0xFFFF) detail::fail("header too long for a version-1 .npy: " + name);
        if (entries_.size() >= 0xFFFE) detail::fail("too many members for the zip32 writer");
        uLong crc = crc32(0L, Z_NULL, 0);
        crc = crc32(crc, hdr.data(), uInt(hdr.size()));
        const unsigned char *p = static_cast<const unsigned char *>(data);
        for (size_t done = 0; done < nbytes;) {
            size_t chunk = std::min<size_t>(nbytes - done, size_t(1) << 30);
            crc = crc32(crc, p + done, uInt(chunk));
            done += chunk;
        }
        Entry e;
        e.name = name + ".npy";
        e.crc = uint32_t(crc);
        e.size = uint32_t(total);
        e.offset = uint64_t(f_.tellp());
        if (e.offset >= 0xFFFFFFFFull) detail::fail("archive too large for zip32 writer");
        put32(0x04034b50u); put16(20); put16(0); put16(0); put16(0); put16(0x21);
        put32(e.crc); put32(e.size); put32(e.size); put16(uint16_t(e.name.size())); put16(0);
        f_.write(e.name.data(), std::streamsize(e.name.size()));
        f_.write(reinterpret_cast<const char *>(hdr.data()), std::streamsize(hdr.size()));
        f_.write(static_cast<const char *>(data), std::streamsize(nbytes));
        entries_.push_back(e);
    }

    /*! \brief Write the central directory and close the file. */
    void close() {
        uint64_t cd_start = uint64_t(f_.tellp());
        for (const Entry &e : entries_) {
            put32(0x02014b50u); put16(20); put16(20); put16(0); put16(0); put16(0); put16(0x21);
            put32(e.crc); put32(e.size); put32(e.size); put16(uint16_t(e.name.size()));
            put16(0); put16(0); put16(0); put16(0); put32(0); put32(uint32_t(e.offset));
            f_.write(e.name.data(), std::streamsize(e.name.size()));
        }
        uint64_t cd_size = uint64_t(f_.tellp()) - cd_start;
        put32(0x06054b50u); put16(0); put16(0); put16(uint16_t(entries_.size())); put16(uint16_t(entries_.size()));
        put32(uint32_t(cd_size)); put32(uint32_t(cd_start)); put16(0);
        f_.close();
    }

private:
    struct Entry { std::string name; uint32_t crc; uint32_t size; uint64_t offset; };
    void put16(uint16_t v) { char b[2] = {char(v & 0xFF), char(v >> 8)}; f_.write(b, 2); }
    void put32(uint32_t v) { char b[4] = {char(v & 0xFF), char((v >> 8) & 0xFF), char((v >> 16) & 0xFF), char(v >> 24)}; f_.write(b, 4); }
    std::ofstream f_;
    std::vector<Entry> entries_;
};

}  // namespace npz
}  // namespace io
}  // namespace graphlily

#endif  // GRAPHLILY_B200_IO_NPZ_H_
