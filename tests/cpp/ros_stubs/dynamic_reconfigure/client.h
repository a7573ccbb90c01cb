#pragma once  // TEST STUB (syntax check only)
#include <string>
namespace dynamic_reconfigure {
template <class C>
class Client {
 public:
  explicit Client(const std::string &);
  bool setConfiguration(const C &);
  bool getCurrentConfiguration(C &);
};
}  // namespace dynamic_reconfigure
