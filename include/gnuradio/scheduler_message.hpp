// gnuradio/scheduler_message.hpp -- wake-up messages exchanged between scheduler threads
// (names follow reference runtime/include/gnuradio/scheduler_message.hpp:8 and
// neighbor_interface.hpp; the implementation is a plain notification record).
#pragma once
#include <memory>

namespace gr {

enum class scheduler_action_t { DONE, NOTIFY_OUTPUT, NOTIFY_INPUT, NOTIFY_ALL, EXIT };

struct scheduler_message {
    scheduler_action_t action;
    explicit scheduler_message(scheduler_action_t a) : action(a) {}
};
typedef std::shared_ptr<scheduler_message> scheduler_message_sptr;
typedef scheduler_message scheduler_action; // reference spelling used at call sites

struct neighbor_interface {
    virtual ~neighbor_interface() {}
    virtual void push_message(scheduler_message_sptr msg) = 0;
};
typedef std::shared_ptr<neighbor_interface> neighbor_interface_sptr;

} // namespace gr
