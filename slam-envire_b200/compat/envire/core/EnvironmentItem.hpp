#pragma once
#include "FrameNode.hpp"
