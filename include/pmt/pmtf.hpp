// pmt/pmtf.hpp -- minimal stand-in for newsched's flatbuffers-backed polymorphic types
// (reference pmt/include/pmt/pmtf.hpp:14-102).  Only what stream tags need: an immutable
// value that is a string, an integer or a double.  The hot path carries raw samples and
// never touches these (pmt is out of scope).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <variant>

namespace pmtf {
class pmt
{
public:
    using value_t = std::variant<std::monostate, int64_t, double, std::string>;
    pmt() = default;
    explicit pmt(value_t v) : _v(std::move(v)) {}
    const value_t& value() const { return _v; }
    bool operator==(const pmt& o) const { return _v == o._v; }

private:
    value_t _v;
};
using pmt_sptr = std::shared_ptr<pmt>;

inline pmt_sptr make_string(const std::string& s) { return std::make_shared<pmt>(pmt::value_t{ s }); }
inline pmt_sptr make_int(int64_t v) { return std::make_shared<pmt>(pmt::value_t{ v }); }
inline pmt_sptr make_double(double v) { return std::make_shared<pmt>(pmt::value_t{ v }); }
inline bool equal(const pmt_sptr& a, const pmt_sptr& b)
{
    if (a == b)
        return true;
    return a && b && *a == *b;
}
} // namespace pmtf
