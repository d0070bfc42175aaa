// Shim (test infrastructure): an in-memory emulation of the slice of the HDF5 C API that the
// reference's EffectiveAction checkpoint code and MeasurementCorrelation writers call
// (e.g. SU2EffectiveAction.hpp:79-205, SU2MeasurementCorrelation.cpp:179-355).
// "Files" live in a process-wide map; the harness serialises them with h5shim::dump_all().
// Lets the UNMODIFIED reference sources compile and run without libhdf5.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <memory>

#ifndef H5SHIM_REAL
#define H5SHIM_REAL float
#endif

typedef int64_t hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
typedef int htri_t;

#define H5E_DEFAULT 0
#define H5P_DEFAULT 0
#define H5S_ALL 0
#define H5F_ACC_RDONLY 0u
#define H5F_ACC_RDWR 1u
#define H5F_ACC_TRUNC 2u
#define H5G_GROUP 0
#define H5G_DATASET 1
#define H5T_NATIVE_FLOAT ((hid_t)-1000)

namespace h5shim {
typedef H5SHIM_REAL real_t;
struct Node {
	bool isGroup = true;
	std::map<std::string, std::shared_ptr<Node>> children; // name-ordered, like HDF5's default index
	std::map<std::string, std::vector<unsigned char>> attributes;
	std::vector<hsize_t> dims;
	size_t elemSize = sizeof(real_t);
	std::vector<unsigned char> data;
};
struct Handle { int kind; std::shared_ptr<Node> node; std::string attr; std::vector<hsize_t> dims; size_t elemSize; };
// kinds: 0 free, 1 file, 2 group, 3 dataset, 4 attribute, 5 dataspace, 6 datatype
inline std::map<std::string, std::shared_ptr<Node>> &files() { static std::map<std::string, std::shared_ptr<Node>> f; return f; }
inline std::vector<Handle> &handles() { static std::vector<Handle> h(1); return h; }
inline hid_t newHandle(const Handle &h) { handles().push_back(h); return (hid_t)handles().size() - 1; }
inline Handle *get(hid_t id) { if (id <= 0 || id >= (hid_t)handles().size() || handles()[id].kind == 0) return nullptr; return &handles()[id]; }
inline herr_t release(hid_t id) { Handle *h = get(id); if (!h) return -1; h->kind = 0; h->node.reset(); return 0; }
inline size_t typeSize(hid_t type) { if (type == H5T_NATIVE_FLOAT) return sizeof(real_t); Handle *h = get(type); return (h && h->kind == 6) ? h->elemSize : sizeof(real_t); }
inline std::shared_ptr<Node> lookup(hid_t loc, const char *name)
{
	Handle *h = get(loc); if (!h || !h->node) return nullptr;
	std::shared_ptr<Node> n = h->node;
	std::string path(name); size_t pos = 0;
	while (pos < path.size())
	{
		size_t slash = path.find('/', pos);
		std::string key = path.substr(pos, slash == std::string::npos ? std::string::npos : slash - pos);
		if (!key.empty()) { auto it = n->children.find(key); if (it == n->children.end()) return nullptr; n = it->second; }
		if (slash == std::string::npos) break;
		pos = slash + 1;
	}
	return n;
}
}

inline herr_t H5Eset_auto(hid_t, void *, void *) { return 0; }
inline htri_t H5Fis_hdf5(const char *name) { return h5shim::files().count(name) ? 1 : -1; }
inline hid_t H5Fopen(const char *name, unsigned, hid_t) { auto it = h5shim::files().find(name); if (it == h5shim::files().end()) return -1; return h5shim::newHandle({ 1, it->second, "", {}, 0 }); }
inline hid_t H5Fcreate(const char *name, unsigned, hid_t, hid_t) { auto n = std::make_shared<h5shim::Node>(); h5shim::files()[name] = n; return h5shim::newHandle({ 1, n, "", {}, 0 }); }
inline herr_t H5Fclose(hid_t id) { return h5shim::release(id); }
inline herr_t H5Gget_num_objs(hid_t loc, hsize_t *num) { h5shim::Handle *h = h5shim::get(loc); if (!h) return -1; *num = h->node->children.size(); return 0; }
inline int H5Gget_objtype_by_idx(hid_t loc, hsize_t idx) { h5shim::Handle *h = h5shim::get(loc); if (!h || idx >= h->node->children.size()) return -1; auto it = h->node->children.begin(); std::advance(it, idx); return it->second->isGroup ? H5G_GROUP : H5G_DATASET; }
inline long H5Gget_objname_by_idx(hid_t loc, hsize_t idx, char *name, size_t size) { h5shim::Handle *h = h5shim::get(loc); if (!h || idx >= h->node->children.size()) return -1; auto it = h->node->children.begin(); std::advance(it, idx); std::strncpy(name, it->first.c_str(), size); if (size) name[size - 1] = 0; return (long)it->first.size(); }
inline hid_t H5Gopen(hid_t loc, const char *name, hid_t) { auto n = h5shim::lookup(loc, name); if (!n || !n->isGroup) return -1; return h5shim::newHandle({ 2, n, "", {}, 0 }); }
inline hid_t H5Gcreate(hid_t loc, const char *name, hid_t, hid_t, hid_t) { h5shim::Handle *h = h5shim::get(loc); if (!h || h->node->children.count(name)) return -1; auto n = std::make_shared<h5shim::Node>(); h->node->children[name] = n; return h5shim::newHandle({ 2, n, "", {}, 0 }); }
inline herr_t H5Gclose(hid_t id) { return h5shim::release(id); }
inline htri_t H5Lexists(hid_t loc, const char *name, hid_t) { return h5shim::lookup(loc, name) ? 1 : 0; }
inline hid_t H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *) { return h5shim::newHandle({ 5, nullptr, "", std::vector<hsize_t>(dims, dims + rank), 0 }); }
inline herr_t H5Sclose(hid_t id) { return h5shim::release(id); }
inline hid_t H5Tarray_create(hid_t base, unsigned rank, const hsize_t *dims) { size_t s = h5shim::typeSize(base); for (unsigned i = 0; i < rank; ++i) s *= dims[i]; return h5shim::newHandle({ 6, nullptr, "", std::vector<hsize_t>(dims, dims + rank), s }); }
inline herr_t H5Tclose(hid_t id) { return h5shim::release(id); }
inline hid_t H5Acreate(hid_t loc, const char *name, hid_t type, hid_t space, hid_t, hid_t)
{
	h5shim::Handle *h = h5shim::get(loc); h5shim::Handle *s = h5shim::get(space); if (!h || !s) return -1;
	size_t n = h5shim::typeSize(type); for (auto d : s->dims) n *= d;
	h->node->attributes[name] = std::vector<unsigned char>(n, 0);
	return h5shim::newHandle({ 4, h->node, name, {}, 0 });
}
inline hid_t H5Aopen(hid_t loc, const char *name, hid_t) { h5shim::Handle *h = h5shim::get(loc); if (!h || !h->node->attributes.count(name)) return -1; return h5shim::newHandle({ 4, h->node, name, {}, 0 }); }
inline herr_t H5Awrite(hid_t attr, hid_t, const void *buf) { h5shim::Handle *h = h5shim::get(attr); if (!h) return -1; auto &a = h->node->attributes[h->attr]; std::memcpy(a.data(), buf, a.size()); return 0; }
inline herr_t H5Aread(hid_t attr, hid_t, void *buf) { h5shim::Handle *h = h5shim::get(attr); if (!h) return -1; auto &a = h->node->attributes[h->attr]; std::memcpy(buf, a.data(), a.size()); return 0; }
inline herr_t H5Aclose(hid_t id) { return h5shim::release(id); }
inline hid_t H5Dcreate(hid_t loc, const char *name, hid_t type, hid_t space, hid_t, hid_t, hid_t)
{
	h5shim::Handle *h = h5shim::get(loc); h5shim::Handle *s = h5shim::get(space); if (!h || !s || h->node->children.count(name)) return -1;
	auto n = std::make_shared<h5shim::Node>(); n->isGroup = false; n->dims = s->dims; n->elemSize = h5shim::typeSize(type);
	size_t bytes = n->elemSize; for (auto d : n->dims) bytes *= d;
	n->data.assign(bytes, 0);
	h->node->children[name] = n;
	return h5shim::newHandle({ 3, n, "", {}, 0 });
}
inline hid_t H5Dopen(hid_t loc, const char *name, hid_t) { auto n = h5shim::lookup(loc, name); if (!n || n->isGroup) return -1; return h5shim::newHandle({ 3, n, "", {}, 0 }); }
inline herr_t H5Dwrite(hid_t ds, hid_t, hid_t, hid_t, hid_t, const void *buf) { h5shim::Handle *h = h5shim::get(ds); if (!h) return -1; std::memcpy(h->node->data.data(), buf, h->node->data.size()); return 0; }
inline herr_t H5Dread(hid_t ds, hid_t, hid_t, hid_t, hid_t, void *buf) { h5shim::Handle *h = h5shim::get(ds); if (!h) return -1; std::memcpy(buf, h->node->data.data(), h->node->data.size()); return 0; }
inline herr_t H5Dclose(hid_t id) { return h5shim::release(id); }
