// Host-side worker threads that cannot take the process down: an exception inside a worker (std::bad_alloc
// while a builder grows its tables, std::system_error from thread creation) is carried back to the caller
// of run() / join() and rethrown there, where the C ABI turns it into a status code (ERT_ERR_NOMEM).
#pragma once
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

namespace ert {

class WorkerGroup {
public:
    WorkerGroup() = default;
    WorkerGroup(const WorkerGroup &) = delete;
    WorkerGroup &operator=(const WorkerGroup &) = delete;
    ~WorkerGroup() { wait(); }

    template <class F>
    void spawn(F f)
    {
        threads_.emplace_back([this, f]() mutable {
            try {
                f();
            } catch (...) {
                std::lock_guard<std::mutex> lock(mu_);
                if (!error_) error_ = std::current_exception();
            }
        });
    }
    // the calling thread's share of the work, under the same rule
    template <class F>
    void inline_run(F f)
    {
        try {
            f();
        } catch (...) {
            std::lock_guard<std::mutex> lock(mu_);
            if (!error_) error_ = std::current_exception();
        }
    }
    void wait()
    {
        for (auto &t : threads_)
            if (t.joinable()) t.join();
        threads_.clear();
    }
    // joins every worker, then rethrows the first exception any of them raised
    void join()
    {
        wait();
        if (error_) {
            std::exception_ptr e = error_;
            error_ = nullptr;
            std::rethrow_exception(e);
        }
    }

private:
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::exception_ptr error_;
};

// f(t) for t in [0, n_thr): t = 0 on the calling thread, the others on workers
template <class F>
void parallel_threads(unsigned n_thr, F f)
{
    WorkerGroup g;
    for (unsigned t = 1; t < n_thr; t++) g.spawn([&f, t] { f(t); });
    g.inline_run([&f] { f(0u); });
    g.join();
}

}  // namespace ert
