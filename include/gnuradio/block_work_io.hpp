// gnuradio/block_work_io.hpp -- work() argument carriers and return codes
// (reference runtime/include/gnuradio/block_work_io.hpp:15-24, :30-39, :45-54).
#pragma once
#include <gnuradio/buffer.hpp>

namespace gr {

struct block_work_input {
    int n_items;
    buffer_sptr buffer;
    int n_consumed; // set by the block; -1 = not set
    block_work_input(int n_items_, buffer_sptr buf) : n_items(n_items_), buffer(std::move(buf)), n_consumed(-1) {}
};

struct block_work_output {
    int n_items;
    buffer_sptr buffer;
    int n_produced; // set by the block; -1 = not set
    block_work_output(int n_items_, buffer_sptr buf) : n_items(n_items_), buffer(std::move(buf)), n_produced(-1) {}
};

enum class work_return_code_t {
    WORK_ERROR = -100,
    WORK_INSUFFICIENT_OUTPUT_ITEMS = -3,
    WORK_INSUFFICIENT_INPUT_ITEMS = -2,
    WORK_DONE = -1,
    WORK_OK = 0,
};

} // namespace gr
