// clusters.out codec: the hps serialisation of cluster_set_t (cluster.hpp:10-42) re-written for flat arrays.
//   vector  -> varint(size) then elements          hps/src/container/vector_serializer.h:15-37
//   int     -> zig-zag then varint                 hps/src/basic_type/int_serializer.h:18-30
//   bool    -> one varint byte 0/1                 hps/src/basic_type/uint_serializer.h:16-43
//   varint  -> little-endian base-128, 0x80 = more hps/src/basic_type/uint_serializer.h:16-32
#include <cstring>
#include <vector>

#include "../../include/rattle_b200.h"

namespace {
inline void put_varint(std::vector<uint8_t> &o, uint64_t v) {
    while (v >= 0x80) {
        o.push_back((uint8_t)(v | 0x80));
        v >>= 7;
    }
    o.push_back((uint8_t)v);
}
inline void put_int(std::vector<uint8_t> &o, int32_t n) { put_varint(o, (uint32_t)((n << 1) ^ (n >> 31))); }
struct Reader {
    const uint8_t *p, *end;
    bool bad = false;
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (true) {
            if (p >= end || shift > 63) {
                bad = true;
                return 0;
            }
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
    }
    int32_t sint() {
        uint32_t z = (uint32_t)varint();
        return (int32_t)((z >> 1) ^ (~(z & 1) + 1));
    }
};
}  // namespace

extern "C" {

int64_t rtl_hps_encode(int n_clusters, const int32_t *main_id, const uint8_t *main_rev, const int32_t *main_gene,
                       const int64_t *cl_off, const int32_t *mem_id, const uint8_t *mem_rev, const int32_t *mem_gene,
                       uint8_t *out, int64_t cap) {
    std::vector<uint8_t> o;
    o.reserve(16 + 6 * (size_t)(cl_off ? cl_off[n_clusters] : 0));
    put_varint(o, (uint64_t)n_clusters);
    for (int c = 0; c < n_clusters; ++c) {
        put_int(o, main_id[c]);
        put_varint(o, main_rev[c] ? 1 : 0);
        put_int(o, main_gene ? main_gene[c] : -1);
        put_varint(o, (uint64_t)(cl_off[c + 1] - cl_off[c]));
        for (int64_t i = cl_off[c]; i < cl_off[c + 1]; ++i) {
            put_int(o, mem_id[i]);
            put_varint(o, mem_rev[i] ? 1 : 0);
            put_int(o, mem_gene ? mem_gene[i] : -1);
        }
    }
    if ((int64_t)o.size() > cap || !out) return -(int64_t)o.size();
    memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}

int rtl_hps_decode(const uint8_t *buf, int64_t len, int32_t *n_clusters, int64_t *n_members, int32_t *main_id,
                   uint8_t *main_rev, int32_t *main_gene, int64_t *cl_off, int32_t *mem_id, uint8_t *mem_rev,
                   int32_t *mem_gene) {
    if (!buf || !n_clusters || !n_members) return RTL_ERR_INPUT;
    Reader r{buf, buf + len};
    const uint64_t nc = r.varint();
    if (r.bad || nc > (uint64_t)len) return RTL_ERR_INPUT;
    const bool fill = main_id != nullptr;
    int64_t o = 0;
    for (uint64_t c = 0; c < nc; ++c) {
        int32_t id = r.sint();
        uint8_t rev = (uint8_t)r.varint();
        int32_t gene = r.sint();
        uint64_t ns = r.varint();
        if (r.bad || ns > (uint64_t)len) return RTL_ERR_INPUT;
        if (fill) {
            main_id[c] = id;
            main_rev[c] = rev;
            if (main_gene) main_gene[c] = gene;
            cl_off[c] = o;
        }
        for (uint64_t i = 0; i < ns; ++i) {
            int32_t mid = r.sint();
            uint8_t mrev = (uint8_t)r.varint();
            int32_t mg = r.sint();
            if (r.bad) return RTL_ERR_INPUT;
            if (fill) {
                mem_id[o] = mid;
                mem_rev[o] = mrev;
                if (mem_gene) mem_gene[o] = mg;
            }
            ++o;
        }
    }
    if (fill) cl_off[nc] = o;
    *n_clusters = (int32_t)nc;
    *n_members = o;
    return RTL_OK;
}
}
