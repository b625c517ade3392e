// No-op archive: the reference's ReadSet::SaveAligments returns immediately (graph.cc:1036) and
// LoadAligments is a no-op when the cache file does not exist (graph.cc:1055). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <ostream>
namespace boost { namespace archive {
struct binary_oarchive { explicit binary_oarchive(std::ostream&) {} template <class T> binary_oarchive& operator<<(const T&) { return *this; } template <class T> binary_oarchive& operator&(const T&) { return *this; } };
}}
