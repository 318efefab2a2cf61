// cli_stubs.cpp -- link-time stand-ins for the three graphapp:: functions the reference CLI's `-md=t` mode calls
// (src/voroUtility.cpp:558-577).  Their real bodies (src/graphapp.cpp) need Boost.Graph, which is not in this image; that
// mode is outside the hot path, so in main_voroUtility_gpu it reports an error instead.
#include <iostream>
#include <voxelcore/graphapp.h>
namespace graphapp
{
bool readGraph(const string&, WeightedGraph&, vector<vector<float>>&)
{
    std::cout << "Error: -md=t needs Boost.Graph, which this build does not have." << std::endl;
    return false;
}
void makeTreeFromGraph(const WeightedGraph&, const vector<vector<float>>&, TreeMethod,
                       vector<NodeHandle>&) {}
void exportTree(const WeightedGraph&, vector<NodeHandle>&, const std::string&) {}
} // namespace graphapp
