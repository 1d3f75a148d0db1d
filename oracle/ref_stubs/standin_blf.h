// TEST INFRASTRUCTURE — BipedalLocomotion's parameter handler and vectors-collection server as far as the MPC sources use
// them: a name -> value table with typed getters (the checker fills it with the numbers of src/config/vs_mcp_config.xml),
// and a logging sink that drops everything.
#pragma once
#include <Eigen/Dense>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <standin_yarp.h>

namespace BipedalLocomotion { namespace ParametersHandler {
class IParametersHandler
{
public:
    typedef std::shared_ptr<IParametersHandler> shared_ptr;
    typedef std::weak_ptr<IParametersHandler> weak_ptr;
    virtual ~IParametersHandler() {}

    bool getParameter(const std::string& n, int& v) const { auto it = m_num.find(n); if (it == m_num.end() || it->second.size() != 1) return false; v = int(it->second[0]); return true; }
    bool getParameter(const std::string& n, double& v) const { auto it = m_num.find(n); if (it == m_num.end() || it->second.size() != 1) return false; v = it->second[0]; return true; }
    bool getParameter(const std::string& n, bool& v) const { auto it = m_num.find(n); if (it == m_num.end() || it->second.size() != 1) return false; v = it->second[0] != 0.0; return true; }
    bool getParameter(const std::string& n, std::string& v) const { auto it = m_str.find(n); if (it == m_str.end() || it->second.size() != 1) return false; v = it->second[0]; return true; }
    bool getParameter(const std::string& n, std::vector<std::string>& v) const { auto it = m_str.find(n); if (it == m_str.end()) return false; v = it->second; return true; }
    bool getParameter(const std::string& n, std::vector<double>& v) const { auto it = m_num.find(n); if (it == m_num.end()) return false; v = it->second; return true; }
    bool getParameter(const std::string& n, std::vector<int>& v) const { auto it = m_num.find(n); if (it == m_num.end()) return false; v.assign(it->second.begin(), it->second.end()); return true; }
    // a dynamically sized vector is resized to the parameter's length; a view must already have it
    template <class S, int R, int C> bool getParameter(const std::string& n, Eigen::Matrix<S, R, C>& v) const
    {
        auto it = m_num.find(n); if (it == m_num.end()) return false;
        v.resize(Eigen::Index(it->second.size()));
        for (size_t i = 0; i < it->second.size(); ++i) v(Eigen::Index(i)) = it->second[i];
        return true;
    }
    bool getParameter(const std::string& n, const Eigen::View& v) const
    {
        auto it = m_num.find(n); if (it == m_num.end() || Eigen::Index(it->second.size()) != v.size()) return false;
        for (size_t i = 0; i < it->second.size(); ++i) v(Eigen::Index(i)) = it->second[i];
        return true;
    }
    weak_ptr getGroup(const std::string& n) const { auto it = m_groups.find(n); return it == m_groups.end() ? weak_ptr() : weak_ptr(it->second); }

    void setNumbers(const std::string& n, const std::vector<double>& v) { m_num[n] = v; }
    void setStrings(const std::string& n, const std::vector<std::string>& v) { m_str[n] = v; }
    void setGroup(const std::string& n, shared_ptr g) { m_groups[n] = g; }
private:
    std::map<std::string, std::vector<double>> m_num;
    std::map<std::string, std::vector<std::string>> m_str;
    std::map<std::string, shared_ptr> m_groups;
};
class YarpImplementation : public IParametersHandler
{
public:
    void set(const yarp::os::Searchable&) {}
};
}} // namespace BipedalLocomotion::ParametersHandler

namespace BipedalLocomotion { namespace YarpUtilities {
class VectorsCollectionServer
{
public:
    bool populateMetadata(const std::string&, const std::vector<std::string>&) { return true; }
    bool finalizeMetadata() { return true; }
    void prepareData() {}
    void clearData() {}
    void sendData(bool = false) {}
    template <class T> bool populateData(const std::string&, const T&) { return true; }
};
}} // namespace BipedalLocomotion::YarpUtilities
