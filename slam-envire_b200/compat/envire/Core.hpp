#pragma once
#include "core/FrameNode.hpp"
#include "maps/Pointcloud.hpp"
#include "../../io/scene.hpp"
