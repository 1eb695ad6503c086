// swift-tools-version:5.5
// UNVERIFIED: no Swift toolchain exists in the build image (DESIGN.md §1).  This package is the thin Swift surface a
// Smelter maintainer would compile against libsmelter_b200.so; it mirrors Package.swift of the reference.
import PackageDescription

let package = Package(
    name: "SmelterB200",
    products: [.library(name: "SmelterB200", targets: ["SmelterB200"])],
    targets: [
        .systemLibrary(name: "CSmelterB200", path: "Sources/CSmelterB200"),
        .target(name: "SmelterB200", dependencies: ["CSmelterB200"], path: "Sources/SmelterB200"),
    ]
)
