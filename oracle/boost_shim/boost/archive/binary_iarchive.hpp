// No-op archive (see binary_oarchive.hpp). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <istream>
namespace boost { namespace archive {
struct binary_iarchive { explicit binary_iarchive(std::istream&) {} template <class T> binary_iarchive& operator>>(T&) { return *this; } template <class T> binary_iarchive& operator&(T&) { return *this; } };
}}
