// Minimal stand-in for boost::archive::binary_{o,i}archive / text_{o,i}archive (TEST ONLY).  Raw little-endian bytes, a
// 64-bit length before strings and vectors; classes go through their serialize(Archive&, unsigned).
#pragma once
#include <cstdint>
#include <cstring>
#include <istream>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>

#include "../serialization/access.hpp"

namespace boost { namespace archive {

class binary_oarchive {
public:
    typedef std::true_type is_saving;
    typedef std::false_type is_loading;
    explicit binary_oarchive(std::ostream& os) : os_(os) {}
    template <class T> binary_oarchive& operator<<(const T& t) { save(t); return *this; }
    template <class T> binary_oarchive& operator&(const T& t) { save(t); return *this; }

private:
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value>::type save(const T& t) {
        os_.write(reinterpret_cast<const char*>(&t), sizeof(T));
    }
    void save(const std::string& s) {
        uint64_t n = s.size();
        save(n);
        os_.write(s.data(), (std::streamsize)n);
    }
    void save(const std::vector<bool>& v) {
        uint64_t n = v.size();
        save(n);
        for (bool b : v) { uint8_t x = b; save(x); }
    }
    template <class T> void save(const std::vector<T>& v) {
        uint64_t n = v.size();
        save(n);
        if constexpr (std::is_arithmetic<T>::value) {
            if (n) os_.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(n * sizeof(T)));
        } else {
            for (const T& e : v) save(e);
        }
    }
    template <class T, size_t N> void save(const T (&a)[N]) {
        for (size_t i = 0; i < N; ++i) save(a[i]);
    }
    template <class T> typename std::enable_if<std::is_class<T>::value>::type save(const T& t) {
        boost::serialization::access::serialize(*this, const_cast<T&>(t), 0u);
    }
    std::ostream& os_;
};

class binary_iarchive {
public:
    typedef std::false_type is_saving;
    typedef std::true_type is_loading;
    explicit binary_iarchive(std::istream& is) : is_(is) {}
    template <class T> binary_iarchive& operator>>(T& t) { load(t); return *this; }
    template <class T> binary_iarchive& operator&(T& t) { load(t); return *this; }

private:
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value>::type load(T& t) {
        is_.read(reinterpret_cast<char*>(&t), sizeof(T));
    }
    void load(std::string& s) {
        uint64_t n = 0;
        load(n);
        s.resize(n);
        if (n) is_.read(&s[0], (std::streamsize)n);
    }
    void load(std::vector<bool>& v) {
        uint64_t n = 0;
        load(n);
        v.resize(n);
        for (uint64_t i = 0; i < n; ++i) { uint8_t x = 0; load(x); v[i] = x != 0; }
    }
    template <class T> void load(std::vector<T>& v) {
        uint64_t n = 0;
        load(n);
        v.resize(n);
        if constexpr (std::is_arithmetic<T>::value) {
            if (n) is_.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
        } else {
            for (T& e : v) load(e);
        }
    }
    template <class T, size_t N> void load(T (&a)[N]) {
        for (size_t i = 0; i < N; ++i) load(a[i]);
    }
    template <class T> typename std::enable_if<std::is_class<T>::value>::type load(T& t) {
        boost::serialization::access::serialize(*this, t, 0u);
    }
    std::istream& is_;
};

typedef binary_oarchive text_oarchive;
typedef binary_iarchive text_iarchive;

}}  // namespace boost::archive
