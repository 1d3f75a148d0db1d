// TEST INFRASTRUCTURE — the five matio calls of utils/src/TrajectoryManager.cpp:61-128 over an in-memory table of named
// FP64 arrays that the checker fills (ref_mpc_add_mat_variable) from the reference's own .mat fixtures.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

enum mat_acc { MAT_ACC_RDONLY = 0, MAT_ACC_RDWR = 1 };
struct matvar_t
{
    char* name;
    void* data;
    size_t* dims;
    std::string name_store;
    std::vector<double> data_store;
    size_t dims_store[2];
};
struct mat_t { std::vector<matvar_t*> vars; size_t next = 0; };
mat_t* standinMatOpen(const char* path);      // defined by the checker (oracle/ref_mpc_shim.cpp): file name -> table
inline mat_t* Mat_Open(const char* path, int) { return standinMatOpen(path); }
inline matvar_t* Mat_VarRead(mat_t* m, const char* name)
{
    for (matvar_t* v : m->vars) if (v->name_store == name) return v;
    return nullptr;
}
inline matvar_t* Mat_VarReadNextInfo(mat_t* m) { return m->next < m->vars.size() ? m->vars[m->next++] : nullptr; }
inline int Mat_Close(mat_t*) { return 0; }
inline void Mat_VarFree(matvar_t*) {}
