#pragma once  // TEST STUB (syntax check only)
namespace std_msgs { struct Float64 { double data = 0; }; }
