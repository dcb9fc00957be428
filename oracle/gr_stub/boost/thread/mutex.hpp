#pragma once
#include <gnuradio/block.h>     // boost::mutex stand-in lives there
