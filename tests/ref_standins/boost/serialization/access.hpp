// Minimal stand-in for Boost.Serialization (TEST ONLY: a real CoGNN build uses Boost).  The reference headers only need
// `boost::serialization::access` as a friend, `ar & member` inside serialize(), and `archive << x` / `archive >> x` over an
// iostream for arithmetic types, std::string, std::vector and classes with serialize().  See binary_oarchive.hpp.
#pragma once
namespace boost { namespace serialization {
class access {
public:
    template <class Archive, class T>
    static void serialize(Archive& ar, T& t, const unsigned int version) { t.serialize(ar, version); }
};
}}  // namespace boost::serialization
#ifndef BOOST_SERIALIZATION_SPLIT_MEMBER
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#endif
