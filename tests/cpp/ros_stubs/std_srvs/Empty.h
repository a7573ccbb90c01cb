#pragma once  // TEST STUB (syntax check only)
namespace std_srvs { struct Empty { struct Request {} request; struct Response {} response; }; }
