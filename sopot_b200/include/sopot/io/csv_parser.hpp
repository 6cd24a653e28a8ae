// io::CsvParser -- host-side table ingestion with the reference's surface (io/csv_parser.hpp:14-132): numeric CSV files
// with '#' comment lines and header cells that do not parse as numbers, in two shapes:
//   parseFile      every numeric cell of every data line (engine, mass and inertia tables: rocket/interpolated_engine.hpp:61-85,
//                  rocket_body.hpp:55-85)
//   parse2DTable   coefficient tables for the bilinear interpolator: the first data line holds the column coordinates, every
//                  later line starts with its row coordinate, which is dropped (:87-131)
// Nothing here runs on the device: the tables a Rocket<T> loads through this parser are packed into one buffer and copied to
// shared memory by the rocket kernel (csrc/rocket.cu).
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sopot::io {

class CsvParser {
public:
    struct Table {
        std::vector<std::vector<double>> data;
        size_t rows() const { return data.size(); }
        size_t cols() const { return data.empty() ? 0 : data.front().size(); }
        std::vector<double> column(size_t col) const {                     // short rows are skipped, as in the reference (:22-31)
            std::vector<double> out;
            out.reserve(data.size());
            for (const auto& r : data)
                if (col < r.size()) out.push_back(r[col]);
            return out;
        }
        const std::vector<double>& row(size_t r) const { return data.at(r); }
        double operator()(size_t r, size_t c) const { return data.at(r).at(c); }
    };

    static Table parseFile(const std::string& filename, char delimiter = ',') { return parse(filename, delimiter, false); }
    static Table parse2DTable(const std::string& filename, char delimiter = ',') { return parse(filename, delimiter, true); }
    // the same two readers on text that is already in memory (additive: tests, embedded tables)
    static Table parseString(const std::string& text, char delimiter = ',', bool table2d = false) {
        std::istringstream in(text);
        return read(in, delimiter, table2d);
    }

private:
    static Table parse(const std::string& filename, char delimiter, bool table2d) {
        std::ifstream file(filename);
        if (!file.is_open()) throw std::runtime_error("Cannot open file: " + filename);      // the reference's text (:46-48)
        return read(file, delimiter, table2d);
    }
    static bool number(std::string cell, double& value) {
        const size_t a = cell.find_first_not_of(" \t\r\n");
        if (a == std::string::npos) return false;
        cell = cell.substr(a, cell.find_last_not_of(" \t\r\n") - a + 1);
        try { value = std::stod(cell); } catch (...) { return false; }                        // a header cell
        return true;
    }
    static Table read(std::istream& in, char delimiter, bool table2d) {
        Table t;
        std::string line, cell;
        bool first_line = true;
        while (std::getline(in, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::vector<double> row;
            std::stringstream ss(line);
            bool first_cell = true;
            while (std::getline(ss, cell, delimiter)) {
                double v;
                // 2-D tables: the leading cell of every line after the first is the row coordinate, not data
                if (number(cell, v) && !(table2d && first_cell && !first_line)) row.push_back(v);
                first_cell = false;
            }
            if (!row.empty()) t.data.push_back(std::move(row));
            first_line = false;
        }
        return t;
    }
};

}  // namespace sopot::io
