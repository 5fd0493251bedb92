#pragma once
#include "stream.hpp"
