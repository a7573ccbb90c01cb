// TEST STUB (syntax check only, see ../README.md): the roscpp surface the planner node and parameter_manager.h use.
#pragma once
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define ROS_INFO(...) ((void)std::printf(__VA_ARGS__))
#define ROS_WARN(...) ((void)std::printf(__VA_ARGS__))
#define ROS_ERROR(...) ((void)std::printf(__VA_ARGS__))
#define ROS_INFO_STREAM(x) (std::cout << x)
#define ROS_WARN_STREAM(x) (std::cout << x)
#define ROS_ERROR_STREAM(x) (std::cout << x)

namespace XmlRpc {
class XmlRpcValue {
 public:
  enum Type { TypeInvalid, TypeBoolean, TypeInt, TypeDouble, TypeString, TypeDateTime, TypeBase64, TypeArray, TypeStruct };
  typedef std::map<std::string, XmlRpcValue> ValueStruct;
  typedef ValueStruct::iterator iterator;
  XmlRpcValue();
  Type getType() const;
  bool hasMember(const std::string &) const;
  int size() const;
  std::string toXml() const;
  XmlRpcValue &operator[](const std::string &);
  const XmlRpcValue &operator[](const std::string &) const;
  XmlRpcValue &operator[](const char *);
  const XmlRpcValue &operator[](const char *) const;
  XmlRpcValue &operator[](int);
  const XmlRpcValue &operator[](int) const;
  operator bool &();
  operator int &();
  operator double &();
  operator std::string &();
  operator const std::string &() const;
  iterator begin();
  iterator end();
};
}  // namespace XmlRpc

namespace ros {
struct Time {
  static Time now();
  double toSec() const;
};
struct Duration {
  Duration(double = 0.0);
  bool sleep() const;
};
struct TimerEvent {};
struct TransportHints {
  TransportHints &reliable();
  TransportHints &tcpNoDelay();
};
struct Publisher {
  template <class M>
  void publish(const M &) const;
};
struct Subscriber {};
struct Timer {
  void start();
  void stop();
};
struct NodeHandle {
  NodeHandle(const std::string & = "");
  template <class M>
  Publisher advertise(const std::string &, unsigned, bool = false);
  template <class M, class T>
  Subscriber subscribe(const std::string &, unsigned, void (T::*)(const M &), T *, const TransportHints & = TransportHints());
  template <class V>
  bool getParam(const std::string &, V &) const;
  template <class V>
  void setParam(const std::string &, const V &) const;
  template <class T>
  Timer createTimer(Duration, void (T::*)(const TimerEvent &), T *, bool = false, bool = true);
};
void spin();
void spinOnce();
void shutdown();
bool ok();
void init(int &, char **, const std::string &);
namespace service {
template <class S>
bool call(const std::string &, S &);
}
}  // namespace ros
