// the one-line replacement of the reference's cf_manager.h
#pragma once
#include <pmaf/cf_manager.hpp>
