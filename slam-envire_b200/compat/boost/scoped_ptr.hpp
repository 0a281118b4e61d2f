// boost::scoped_ptr stand-in (used by the reference's icp/icpLocalization.hpp:12-13,84): owning pointer, swap, no copy
#pragma once
namespace boost {
template <class T> class scoped_ptr {
public:
    explicit scoped_ptr(T *p = 0) : p_(p) {}
    ~scoped_ptr() { delete p_; }
    T *get() const { return p_; }
    T *operator->() const { return p_; }
    T &operator*() const { return *p_; }
    void reset(T *p = 0) { if (p != p_) { delete p_; p_ = p; } }
    void swap(scoped_ptr &o) { T *t = p_; p_ = o.p_; o.p_ = t; }
    void swap(scoped_ptr &&o) { T *t = p_; p_ = o.p_; o.p_ = t; }
    explicit operator bool() const { return p_ != 0; }
private:
    scoped_ptr(const scoped_ptr &);
    scoped_ptr &operator=(const scoped_ptr &);
    T *p_;
};
} // namespace boost
