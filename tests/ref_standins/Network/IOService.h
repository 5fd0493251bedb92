#pragma once
#include "../Common/Defines.h"
