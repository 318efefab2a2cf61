// Type-only stand-in for <boost/graph/undirected_graph.hpp>.
//
// include/graphapp.h:4,24-29 of the reference names boost::undirected_graph and graph_traits in
// its declarations; Boost is fetched by Bazel (WORKSPACE:28-49) and is absent from this image.
// The `-md=t` skeleton-tree mode that needs the real library is out of scope and
// is stubbed to an error at link time; this header only lets the CLI translation unit compile.
#pragma once
#include <cstddef>
namespace boost
{
namespace graph {}
template <class VP, class EP> struct undirected_graph
{
    typedef std::size_t vertex_descriptor;
    typedef std::size_t edge_descriptor;
    typedef const std::size_t* vertex_iterator;
    typedef const std::size_t* edge_iterator;
};
template <class G> struct graph_traits
{
    typedef typename G::vertex_descriptor vertex_descriptor;
    typedef typename G::edge_descriptor edge_descriptor;
    typedef typename G::vertex_iterator vertex_iterator;
    typedef typename G::edge_iterator edge_iterator;
};
} // namespace boost
