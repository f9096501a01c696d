/* shim: boost::mpi::communicator for an image without MPI (test infrastructure only).
 *
 * Two modes.
 *   - No world set up (the default): rank 0 of 1, every data-moving member aborts.  This is all that
 *     JobTree::run_master(boost::mpi::communicator&, ...) (master.hh:121, master.cc:551-556) and worker<>()
 *     (worker.hh:53-177) need to type-check when the oracle drives the reference tree through the Communicator
 *     *concept* (masterloop.cc:21-22), exactly like testmaster.cc does.
 *   - boost::mpi::shim::World set up by `pf_ref mpirun` (oracle/ref_driver.cc): every RANK IS A THREAD of this
 *     process, and send / recv / isend / iprobe move copies of the message objects through per-rank mailboxes
 *     (mutex + condition variable) with MPI's matching rules: (source, tag) with wildcards, non-overtaking per
 *     sender.  The reference's master (masterloop.cc) and worker (worker.hh) loops run on it UNMODIFIED -- the
 *     stock `mpirun -np N ./fetching` minus the wire: what the CPU baseline times.
 */
#ifndef PF_SHIM_BOOST_MPI_COMMUNICATOR_HPP
#define PF_SHIM_BOOST_MPI_COMMUNICATOR_HPP
#include <stdlib.h>

#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <vector>

#include "../optional.hpp"
#include "datatype.hpp"
#ifndef MPI_ANY_TAG
#define MPI_ANY_TAG (-1)
#endif
#ifndef MPI_ANY_SOURCE
#define MPI_ANY_SOURCE (-2)
#endif
namespace boost { namespace mpi {
namespace shim {
struct Message {
    int source, tag;
    std::shared_ptr<void> payload;  // a copy of the object sent (nullptr for an empty message)
};
struct Mailbox {
    std::mutex m;
    std::condition_variable cv;
    std::deque<Message> q;
};
struct World {
    explicit World(int n) : size(n), box(n) {
        for (auto& b : box) b.reset(new Mailbox);
    }
    int size;
    std::vector<std::unique_ptr<Mailbox> > box;
    std::mutex bm;
    std::condition_variable bcv;
    int barrier_count = 0, barrier_round = 0;
};
inline World*& world() {
    static World* w = nullptr;
    return w;
}
inline int& my_rank() {
    static thread_local int r = 0;
    return r;
}
inline bool matches(const Message& msg, int source, int tag) {
    return (source == MPI_ANY_SOURCE || msg.source == source) && (tag == MPI_ANY_TAG || msg.tag == tag);
}
}  // namespace shim

class status {
  public:
    status(int source = 0, int tag = 0) : source_(source), tag_(tag) {}
    int source() const { return source_; }
    int tag() const { return tag_; }

  private:
    int source_, tag_;
};
class request {
  public:
    bool test() { return true; }  // a send is complete as soon as the copy sits in the mailbox
};
class communicator {
  public:
    int rank() const { return shim::world() ? shim::my_rank() : 0; }
    int size() const { return shim::world() ? shim::world()->size : 1; }
    void barrier() const {
        shim::World* w = shim::world();
        if (!w) return;
        std::unique_lock<std::mutex> l(w->bm);
        const int round = w->barrier_round;
        if (++w->barrier_count == w->size) {
            w->barrier_count = 0;
            ++w->barrier_round;
            w->bcv.notify_all();
        } else
            w->bcv.wait(l, [&] { return w->barrier_round != round; });
    }
    boost::optional<status> iprobe(int source, int tag) const {
        shim::Mailbox& b = mine();
        std::lock_guard<std::mutex> l(b.m);
        for (const shim::Message& msg : b.q)
            if (shim::matches(msg, source, tag)) return boost::optional<status>(status(msg.source, msg.tag));
        return boost::optional<status>();
    }
    status recv(int source, int tag) const {
        shim::Message msg = take(source, tag);
        return status(msg.source, msg.tag);
    }
    template <typename T> status recv(int source, int tag, T& value) const {
        shim::Message msg = take(source, tag);
        if (!msg.payload) abort();  // an empty message where an object was expected
        value = *static_cast<const T*>(msg.payload.get());
        return status(msg.source, msg.tag);
    }
    void send(int dest, int tag) const { post(dest, tag, std::shared_ptr<void>()); }
    template <typename T> void send(int dest, int tag, const T& value) const {
        post(dest, tag, std::shared_ptr<void>(new T(value), [](void* p) { delete static_cast<T*>(p); }));
    }
    template <typename T> request isend(int dest, int tag, const T& value) const {
        send(dest, tag, value);
        return request();
    }

  private:
    shim::Mailbox& mine() const {
        shim::World* w = shim::world();
        if (!w) abort();  // no MPI in this image and no thread world set up
        return *w->box[shim::my_rank()];
    }
    void post(int dest, int tag, std::shared_ptr<void> payload) const {
        shim::World* w = shim::world();
        if (!w || dest < 0 || dest >= w->size) abort();
        shim::Mailbox& b = *w->box[dest];
        {
            std::lock_guard<std::mutex> l(b.m);
            b.q.push_back(shim::Message{shim::my_rank(), tag, std::move(payload)});
        }
        b.cv.notify_all();
    }
    shim::Message take(int source, int tag) const {
        shim::Mailbox& b = mine();
        std::unique_lock<std::mutex> l(b.m);
        for (;;) {
            for (auto it = b.q.begin(); it != b.q.end(); ++it)
                if (shim::matches(*it, source, tag)) {
                    shim::Message msg = std::move(*it);
                    b.q.erase(it);
                    return msg;
                }
            b.cv.wait(l);
        }
    }
};
} }
#endif
