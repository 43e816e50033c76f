// gnuradio/tag.hpp -- stream tags and propagation policies
// (same names and meaning as reference runtime/include/gnuradio/tag.hpp:8-43).
#pragma once
#include <pmt/pmtf.hpp>
#include <cstdint>

namespace gr {

enum class tag_propagation_policy_t { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2, TPP_CUSTOM = 3 };

class tag_t
{
public:
    uint64_t offset; // absolute item offset in the stream
    pmtf::pmt_sptr key, value, srcid;
    tag_t(uint64_t offset_, pmtf::pmt_sptr key_, pmtf::pmt_sptr value_, pmtf::pmt_sptr srcid_ = nullptr)
        : offset(offset_), key(std::move(key_)), value(std::move(value_)), srcid(std::move(srcid_))
    {
    }
    bool operator==(const tag_t& r) const
    {
        return offset == r.offset && pmtf::equal(key, r.key) && pmtf::equal(value, r.value);
    }
    bool operator!=(const tag_t& r) const { return !(*this == r); }
};

} // namespace gr
