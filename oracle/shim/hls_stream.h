// hls_stream.h -- stand-in for the Vitis HLS stream header: an unbounded FIFO.
// TEST INFRASTRUCTURE ONLY (see ap_int.h).  The reference's dataflow stages run one after the other in
// software, each draining the FIFO the previous one filled, which is the C-simulation semantics of
// `#pragma HLS dataflow`.
#pragma once

#include <deque>

namespace hls {
template <typename T>
class stream {
    std::deque<T> q;

public:
    stream() {}
    explicit stream(const char *) {}
    stream(const stream &) = delete;
    bool empty() const { return q.empty(); }
    size_t size() const { return q.size(); }
    void write(const T &v) { q.push_back(v); }
    T read() {
        T v = q.front();
        q.pop_front();
        return v;
    }
    void read(T &v) { v = read(); }
    stream &operator<<(const T &v) { write(v); return *this; }
    stream &operator>>(T &v) { v = read(); return *this; }
};
}  // namespace hls
