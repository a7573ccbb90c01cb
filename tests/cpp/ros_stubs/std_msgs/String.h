#pragma once  // TEST STUB (syntax check only)
#include <string>
namespace std_msgs { struct String { std::string data; }; }
