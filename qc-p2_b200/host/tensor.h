// tensor.h -- minimal host containers standing in for the reference's third-party types
// (general.h:24-26, hf.h:26-27): DTensor = btas::Tensor<double> (dense, row-major, zero
// initialised), Matrix = Eigen row-major dynamic matrix, Vector = column vector.  Only the
// surface the reference actually uses is provided.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <vector>

typedef double real_t;

class DTensor {
public:
    DTensor() {}
    template<class... I>
    explicit DTensor(I... e) : ext_{static_cast<long>(e)...} {
        long n = 1;
        for (long x : ext_) n *= x;
        v_.assign(static_cast<size_t>(n), 0.0);
    }
    int rank() const { return static_cast<int>(ext_.size()); }
    long extent(int k) const { return ext_[static_cast<size_t>(k)]; }
    size_t size() const { return v_.size(); }
    bool empty() const { return v_.empty(); }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    void fill(double x) { v_.assign(v_.size(), x); }
    template<class... I>
    double &operator()(I... idx) { return v_[offset(idx...)]; }
    template<class... I>
    double operator()(I... idx) const { return v_[offset(idx...)]; }
    DTensor &operator+=(const DTensor &o) {
        if (o.v_.size() != v_.size()) throw std::invalid_argument("DTensor +=: size mismatch");
        for (size_t k = 0; k < v_.size(); ++k) v_[k] += o.v_[k];
        return *this;
    }
    DTensor &operator-=(const DTensor &o) {
        if (o.v_.size() != v_.size()) throw std::invalid_argument("DTensor -=: size mismatch");
        for (size_t k = 0; k < v_.size(); ++k) v_[k] -= o.v_[k];
        return *this;
    }

private:
    template<class... I>
    size_t offset(I... idx) const {
        const long ix[] = {static_cast<long>(idx)...};
        assert(sizeof...(I) == ext_.size());
        size_t o = 0;
        for (size_t k = 0; k < sizeof...(I); ++k) o = o * static_cast<size_t>(ext_[k]) + static_cast<size_t>(ix[k]);
        return o;
    }
    std::vector<long> ext_;
    std::vector<double> v_;
};

class Matrix {
public:
    Matrix() {}
    Matrix(long r, long c) : r_(r), c_(c), v_(static_cast<size_t>(r * c), 0.0) {}
    static Matrix Zero(long r, long c) { return Matrix(r, c); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    double &operator()(long i, long j) { return v_[static_cast<size_t>(i * c_ + j)]; }
    double operator()(long i, long j) const { return v_[static_cast<size_t>(i * c_ + j)]; }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    Matrix transpose() const {
        Matrix t(c_, r_);
        for (long i = 0; i < r_; ++i)
            for (long j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    Matrix operator*(const Matrix &o) const {
        if (c_ != o.r_) throw std::invalid_argument("Matrix *: shape mismatch");
        Matrix p(r_, o.c_);
        for (long i = 0; i < r_; ++i)
            for (long k = 0; k < c_; ++k) {
                const double a = (*this)(i, k);
                for (long j = 0; j < o.c_; ++j) p(i, j) += a * o(k, j);
            }
        return p;
    }
    Matrix operator+(const Matrix &o) const {
        Matrix s(*this);
        for (size_t k = 0; k < v_.size(); ++k) s.v_[k] += o.v_[k];
        return s;
    }
    Matrix operator-(const Matrix &o) const {
        Matrix s(*this);
        for (size_t k = 0; k < v_.size(); ++k) s.v_[k] -= o.v_[k];
        return s;
    }
    Matrix &operator+=(const Matrix &o) {
        for (size_t k = 0; k < v_.size(); ++k) v_[k] += o.v_[k];
        return *this;
    }
    Matrix operator*(double s) const {
        Matrix m(*this);
        for (double &x : m.v_) x *= s;
        return m;
    }
    Matrix leftCols(long k) const {
        Matrix m(r_, k);
        for (long i = 0; i < r_; ++i)
            for (long j = 0; j < k; ++j) m(i, j) = (*this)(i, j);
        return m;
    }
    double norm() const {// Frobenius
        double s = 0.0;
        for (double x : v_) s += x * x;
        return std::sqrt(s);
    }

private:
    long r_ = 0, c_ = 0;
    std::vector<double> v_;
};

class Vector {
public:
    Vector() {}
    explicit Vector(long n) : v_(static_cast<size_t>(n), 0.0) {}
    long size() const { return static_cast<long>(v_.size()); }
    double &operator()(long i) { return v_[static_cast<size_t>(i)]; }
    double operator()(long i) const { return v_[static_cast<size_t>(i)]; }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }

private:
    std::vector<double> v_;
};
