// Minimal stand-in for the Boost.Iostreams pieces comm_sync.h uses (TEST ONLY): a stream that appends to a std::string
// (back_insert_device) and one that reads from a char array (basic_array_source).
#pragma once
#include <istream>
#include <ostream>
#include <streambuf>
#include <string>

namespace boost { namespace iostreams {

template <class Container>
class back_insert_device {
public:
    explicit back_insert_device(Container& c) : c_(&c) {}
    Container* c_;
};

template <class Ch>
class basic_array_source {
public:
    basic_array_source(const Ch* p, size_t n) : p_(p), n_(n) {}
    const Ch* p_;
    size_t n_;
};

template <class Device> class stream;

template <class Container>
class stream<back_insert_device<Container>> : public std::ostream {
    struct buf : std::streambuf {
        Container* c;
        int_type overflow(int_type ch) override {
            if (ch != traits_type::eof()) c->push_back((char)ch);
            return ch;
        }
        std::streamsize xsputn(const char* s, std::streamsize n) override {
            c->append(s, (size_t)n);
            return n;
        }
    } b_;

public:
    explicit stream(back_insert_device<Container>& d) : std::ostream(nullptr) {
        b_.c = d.c_;
        rdbuf(&b_);
    }
};

template <class Ch>
class stream<basic_array_source<Ch>> : public std::istream {
    struct buf : std::streambuf {
        buf(const Ch* p, size_t n) {
            char* q = const_cast<char*>(reinterpret_cast<const char*>(p));
            setg(q, q, q + n);
        }
    } b_;

public:
    explicit stream(basic_array_source<Ch>& d) : std::istream(nullptr), b_(d.p_, d.n_) { rdbuf(&b_); }
};

}}  // namespace boost::iostreams
