// TEST INFRASTRUCTURE — the YARP names the MPC sources mention: log streams (errors go to stderr, the rest is dropped)
// and the configuration-file classes that utils/src/FlightControlUtils.cpp's XML reader is written against (declared so
// that the file compiles; the checker never reads XML — parameters are set directly on the handler stand-in).
#pragma once
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace yarp { namespace os {
class LogStream : public std::runtime_error
{
public:
    explicit LogStream(bool print) : std::runtime_error("yarp log stream (stand-in)"), m_print(print) {}
    LogStream(const LogStream& o) : std::runtime_error(o), m_print(false) {}
    ~LogStream() { if (m_print) std::cerr << std::endl; }
    template <class T> LogStream& operator<<(const T& v) { if (m_print) std::cerr << v << " "; return *this; }
private:
    bool m_print;
};
class Searchable {};
class Bottle
{
public:
    Bottle& addList() { return *this; }
    void addString(const std::string&) {}
};
class Property : public Searchable
{
public:
    void addGroup(const std::string&) {}
    Bottle& findGroup(const std::string&) { return m_b; }
    void fromString(const std::string&, bool = true) {}
private:
    Bottle m_b;
};
class ResourceFinder
{
public:
    std::string findFileByName(const std::string& name) { return name; }
};
}} // namespace yarp::os
inline yarp::os::LogStream yError() { std::cerr << "[reference yError] "; return yarp::os::LogStream(true); }
inline yarp::os::LogStream yWarning() { return yarp::os::LogStream(false); }
inline yarp::os::LogStream yInfo() { return yarp::os::LogStream(false); }
inline yarp::os::LogStream yDebug() { return yarp::os::LogStream(false); }

namespace yarp { namespace robotinterface {
class Param
{
public:
    std::string name() const { return m_name; }
    std::string value() const { return m_value; }
    std::string m_name, m_value;
};
typedef std::vector<Param> ParamList;
inline ParamList mergeDuplicateGroups(const ParamList& p) { return p; }
class Device
{
public:
    ParamList params() const { return ParamList(); }
};
class Robot
{
public:
    Device device(const std::string&) const { return Device(); }
};
struct XMLReaderResult { bool parsingIsSuccessful = false; Robot robot; };
class XMLReader
{
public:
    XMLReaderResult getRobotFromFile(const std::string&) { return XMLReaderResult(); }
};
}} // namespace yarp::robotinterface
namespace yarp { namespace sig { class Vector {}; }}
