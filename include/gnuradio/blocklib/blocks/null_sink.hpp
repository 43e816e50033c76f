// gnuradio/blocklib/blocks/null_sink.hpp -- include-path compatibility with the reference tree;
// the harness blocks live together in host_blocks.hpp.
#pragma once
#include <gnuradio/blocklib/blocks/host_blocks.hpp>
