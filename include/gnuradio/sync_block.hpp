// gnuradio/sync_block.hpp -- 1:1 rate block wrapper.
// Contract of reference runtime/include/gnuradio/sync_block.hpp:36-86: clamp every port to the
// minimum n_items, call work(), require equal n_produced on all outputs, set n_consumed =
// n_produced on every input.
#pragma once
#include <gnuradio/block.hpp>

#include <limits>

namespace gr {

class sync_block : public block
{
public:
    explicit sync_block(const std::string& name) : block(name) {}

    work_return_code_t do_work(std::vector<block_work_input>& work_input,
                               std::vector<block_work_output>& work_output) override
    {
        int n = std::numeric_limits<int>::max();
        for (auto& w : work_input)
            n = std::min(n, w.n_items);
        for (auto& w : work_output)
            n = std::min(n, w.n_items);
        for (auto& w : work_input)
            w.n_items = n;
        for (auto& w : work_output)
            w.n_items = n;

        work_return_code_t ret = work(work_input, work_output);

        int produced = -1;
        for (size_t i = 0; i < work_output.size(); i++) {
            if (i == 0)
                produced = work_output[i].n_produced;
            else if (work_output[i].n_produced != produced)
                throw std::runtime_error("outputs for sync_block must produce same number of items");
        }
        for (auto& w : work_input)
            w.n_consumed = produced < 0 ? w.n_items : produced;
        return ret;
    }
};

} // namespace gr
