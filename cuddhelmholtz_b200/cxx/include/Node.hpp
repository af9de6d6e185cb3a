// Mesh node record (reference include/Node.hpp:8-27).
#ifndef CUDDH_NODES_HPP
#define CUDDH_NODES_HPP

#include <vector>

namespace cuddh
{
    enum class NodeType { INTERIOR, BOUNDARY };

    struct Node
    {
        struct element_info
        {
            int i;  ///< corner of the element this node is (0..3)
            int id; ///< global element index
        };
        int id;
        NodeType type;
        double x[2];
        std::vector<element_info> connected_elements; ///< in element order
    };
} // namespace cuddh

#endif
