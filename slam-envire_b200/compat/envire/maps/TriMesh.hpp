// envire-mini: TriMesh is only used for its vertex-attribute constants on this path (typedef in Pointcloud.hpp)
#pragma once
#include "Pointcloud.hpp"
