// blocks::null_sink -- discards its input (reference blocklib/blocks/include/gnuradio/blocklib/
// blocks/null_sink.hpp:41-45).  Never touches the data, so it works on host and device edges.
#pragma once
#include <gnuradio/sync_block.hpp>

namespace gr {
namespace blocks {

class null_sink : public sync_block
{
public:
    typedef std::shared_ptr<null_sink> sptr;
    static sptr make(size_t itemsize, size_t nports = 1)
    {
        auto ptr = std::make_shared<null_sink>(itemsize, nports);
        for (size_t i = 0; i < nports; i++)
            ptr->add_port(untyped_port::make("in" + std::to_string(i), port_direction_t::INPUT, itemsize));
        return ptr;
    }
    null_sink(size_t itemsize, size_t nports) : sync_block("null_sink"), _itemsize(itemsize), _nports(nports) {}
    work_return_code_t work(std::vector<block_work_input>& work_input, std::vector<block_work_output>&) override
    {
        for (auto& w : work_input) {
            w.n_consumed = w.n_items;
            _n_items += w.n_items;
        }
        return work_return_code_t::WORK_OK;
    }
    uint64_t n_items() const { return _n_items; }

private:
    size_t _itemsize, _nports;
    uint64_t _n_items = 0;
};

} // namespace blocks
} // namespace gr
