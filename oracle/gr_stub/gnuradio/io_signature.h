#pragma once
#include "block.h"
